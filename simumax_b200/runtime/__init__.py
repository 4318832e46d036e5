"""Measured side of predicted-vs-measured: training steps built only from this repo's sm_100a kernels.

``LlamaTrainer`` (data parallel, ZeRO-1), ``TPLlamaTrainer`` (tensor + sequence parallel on the fused AG->GEMM / GEMM->RS
kernels), ``PPLlamaTrainer`` (1F1B pipeline), ``EPMoELayer`` (expert-parallel MoE block).  Imported lazily: they need CUDA.
"""


def __getattr__(name):
    if name in ("LlamaTrainer", "LlamaDims"):
        from . import trainer
        return getattr(trainer, name)
    if name == "TPLlamaTrainer":
        from .tp_trainer import TPLlamaTrainer
        return TPLlamaTrainer
    if name == "PPLlamaTrainer":
        from .pp_trainer import PPLlamaTrainer
        return PPLlamaTrainer
    if name == "EPMoELayer":
        from .moe_layer import EPMoELayer
        return EPMoELayer
    raise AttributeError(name)
