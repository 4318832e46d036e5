"""PerfLLM: the orchestrator / public facade.

    perf = PerfLLM()
    perf.configure(strategy_config=..., model_config=..., system_config=...)   # paths or config objects
    perf.run_estimate()                 # build stage chunks + evaluate every op
    mem  = perf.analysis_mem()          # per-stage peak memory
    cost = perf.analysis_cost()         # iteration time, TGS, TFLOPS, MFU, breakdowns
    perf.analysis(save_path)            # both + result files     (perf.perf() = run_estimate + analysis)
    perf.simulate(save_path)            # discrete-event replay -> Chrome trace + memory snapshot
    perf.search_*()                     # micro-batch / recompute / parallel-strategy search

Method names, result keys and output files match the reference (simumax/core/perf_llm.py) so scripts
port unchanged; internally everything runs on the flat op IR (core/ir.py, models/llm.py).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Union

from .config import (ENABLE_SIMU_GRAPH, SIMU_CHECK, TMP_PATH, ModelConfig, StrategyConfig, SystemConfig,
                     set_capture_graph_only)
from .records import PeakPoint, Result
from .schedule import (effective_straggler_samples, one_f_one_b, straggler_increase_ratio)
from .utils import HumanReadableSize, convert_final_result_to_human_format, get_pp_p2p_comm_size
from ..models.llm import StageChunk

FIRST_CHUNK = "first_stage_chunk"
MIDDLE_CHUNK = "middle_stage_chunk"
LAST_CHUNK = "last_stage_chunk"
_STAGE_NAME = {FIRST_CHUNK: "first", MIDDLE_CHUNK: "middle", LAST_CHUNK: "last"}


def get_effective_straggler_sample_count(world_size, num_per_node, dp_size, edp_size) -> int:
    return effective_straggler_samples(world_size, num_per_node, dp_size, edp_size)


def estimate_straggler_increase_ratio(worker_count: int) -> float:
    return straggler_increase_ratio(worker_count)


class PerfBase:
    dtype_to_element_size = {"fp32": 4, "fp16": 2, "bf16": 2}

    def __init__(self) -> None:
        self.is_configured = False
        self.strategy: StrategyConfig = None
        self.model_config: ModelConfig = None
        self.system: SystemConfig = None
        self.graph = None
        self.debug_points: List[str] = []
        self.debug_points_last_stage: List[str] = []

    # ---------------------------------------------------------------- configuration
    def configure(self, strategy_config: Union[StrategyConfig, str] = None,
                  model_config: Union[ModelConfig, str] = None, system_config: Union[SystemConfig, str] = None,
                  debug_points: List[str] = None, debug_points_last_stage=None):
        if not isinstance(strategy_config, StrategyConfig):
            strategy_config = StrategyConfig.init_from_config_file(strategy_config)
        strategy_config.sanity_check()
        self.strategy = strategy_config
        if not isinstance(model_config, ModelConfig):
            model_config = ModelConfig.init_from_config_file(model_config)
        model_config.sanity_check()
        self.model_config = model_config
        if not isinstance(system_config, SystemConfig):
            system_config = SystemConfig.init_from_config_file(system_config)
        system_config.sanity_check()
        self.system = system_config
        self.debug_points = debug_points if debug_points is not None else []
        self.debug_points_last_stage = debug_points_last_stage if debug_points_last_stage is not None else []
        self._cross_sanity_check()
        self.is_configured = True

    def _cross_sanity_check(self):
        raise NotImplementedError

    # ---------------------------------------------------------------- network class selection
    def analysis_net(self, re_analysis: bool = False):
        """Pick the network class of every parallel dimension from how many ranks its groups span.

        NVLink machines: group fits in a node -> ``high_intra_node`` else ``inter_node``.
        PCIe machines: ``intra_node_pcie_{2,4,8}x`` tiers.  (reference perf_llm.py:369-474)
        """
        s, npn = self.strategy, self.system.num_per_node
        tp, cp, etp, edp, ep, pp, dp = s.tp_size, s.cp_size, s.etp_size, s.edp_size, s.ep_size, s.pp_size, s.dp_size
        if self.system.intra_with_pcie:
            def pick(span):
                if span <= 2:
                    return "intra_node_pcie_2x"
                if span <= 4:
                    return "intra_node_pcie_4x"
                if span <= 8:
                    return "intra_node_pcie_8x"
                return "inter_node"
            choice = {"pp_net": pick(tp * dp * pp * cp), "ep_net": pick(ep * etp), "tp_net": pick(tp),
                      "cp_net": pick(tp * cp), "etp_net": pick(etp), "dp_net": pick(tp * cp * dp),
                      "edp_net": pick(etp * ep * edp)}
        else:
            def pick(fits):
                return "high_intra_node" if fits else "inter_node"
            choice = {"pp_net": pick(s.world_size // pp < npn), "ep_net": pick(ep * etp <= npn), "tp_net": pick(tp <= npn),
                      "cp_net": pick(tp * cp <= npn), "etp_net": pick(etp <= npn), "dp_net": pick(tp * cp * dp <= npn),
                      "edp_net": pick(etp * ep * edp <= npn)}
        for attr, value in choice.items():
            if getattr(s, attr) == "auto" or re_analysis:
                setattr(s, attr, value)

    def run_estimate(self, capture_graph: bool = False, save_path: str = "./"):
        assert self.is_configured, "should call configure() first"
        self.model_config.maybe_pad_vocab_size(self.strategy.tp_size, log=getattr(self, "_search_verbose", True))
        self.analysis_net(re_analysis=True)
        self.build()
        if capture_graph:
            self.graph = self.capture(save_path)
        self._run()
        self._write_debug_points()

    def _write_debug_points(self) -> None:
        """``debug_points`` (first stage) / ``debug_points_last_stage``: ops whose scope path (``"GPTModel -> ... -> LinearCol"``)
        or dotted name is listed get their F / B / W / recompute / collective times appended to ``<TMP_PATH>/cost_log.json``
        (reference base_struct.py:771-800)."""
        wanted = [(FIRST_CHUNK, self.debug_points or []), (LAST_CHUNK, self.debug_points_last_stage or [])]
        if not any(points for _, points in wanted):
            return
        found = {}
        for key, points in wanted:
            chunks = getattr(self, "model_chunk_dict", {})
            chunk = chunks.get(key) or (chunks.get(FIRST_CHUNK) if key == LAST_CHUNK else None)      # pp = 1: one chunk is both
            if chunk is None or not points:
                continue
            for n in chunk.nodes:
                hit = next((pt for pt in points if pt in (n.path, n.name) or n.path.endswith(pt) or n.name.endswith(pt)), None)
                if hit is None:
                    continue
                c = n.cost
                found[f"{key}: {n.path} ({n.name})"] = {"cost_F": c.fwd_compute_time, "cost_B": c.bwd_grad_act_time,
                                                          "cost_W": c.bwd_grad_w_time, "recompute_F": c.recompute_compute_time,
                                                          "net_F": c.fwd_net_time, "net_B": c.bwd_grad_act_net_time + c.bwd_grad_w_net_time}
        if not found:
            return
        os.makedirs(TMP_PATH, exist_ok=True)
        path = os.path.join(TMP_PATH, "cost_log.json")
        try:
            existing = json.load(open(path, encoding="utf-8")) if os.path.exists(path) else {}
        except json.JSONDecodeError:
            existing = {}
        existing.update(found)
        with open(path, "w", encoding="utf-8") as fh:
            json.dump(existing, fh, indent=4, ensure_ascii=False)


class PerfLLM(PerfBase):
    """Performance / memory model of one training iteration of a decoder-only LLM."""

    def __init__(self) -> None:
        super().__init__()
        self.model_chunk_dict: Dict[str, StageChunk] = {}
        self.vpp_chunk_dict: Dict[str, StageChunk] = {}
        self.vpp_stage_chunk_names: Dict[str, List[str]] = {}
        self.pp_state_peak_point: Dict[str, PeakPoint] = {}
        self.enable_chunk_profile_cache = False
        self.enable_search_unit_profile_cache = False
        self._search_verbose = True

    # ---------------------------------------------------------------- validation
    def _cross_sanity_check(self):
        s, m = self.strategy, self.model_config
        assert self.debug_points is None or isinstance(self.debug_points, list), "debug_points should be a list"
        if s.megatron_recompute:
            mods = s.megatron_recompute_module_set
            if "mla_up_proj" in mods:
                assert getattr(m, "attention_type", None) == "mla", \
                    "megatron_recompute mla_up_proj is only supported with MLA attention"
            if "moe_act" in mods:
                assert m.expert_num > 1, "megatron_recompute moe_act is only supported on MoE models"
                assert m.group_linear_mode == "parallel", \
                    "megatron_recompute moe_act is only supported with grouped-gemm MoE"
            if s.fp8:
                assert not (mods & {"layernorm", "moe_act"}), \
                    "megatron_recompute layernorm and moe_act cannot work with fp8"
        assert m.head_num % s.tp_size == 0, f"head_num {m.head_num} should be divisible by tp_size {s.tp_size}"
        if m.kv_head_num is not None:
            assert m.kv_head_num % s.tp_size == 0, \
                f"kv_head_num {m.kv_head_num} should be divisible by tp_size {s.tp_size}"
        assert m.expert_num % s.ep_size == 0, f"expert num {m.expert_num} should be divisible by ep_size {s.ep_size}"
        if s.cp_size > 1 and s.cp_comm_type == "a2a":
            assert m.head_num % s.cp_size == 0, \
                f"head_num {m.head_num} must be divisible by cp_size {s.cp_size} when cp_comm_type='a2a'"
            if m.kv_head_num is not None:
                assert m.kv_head_num % s.cp_size == 0, \
                    f"kv_head_num {m.kv_head_num} must be divisible by cp_size {s.cp_size} when cp_comm_type='a2a'"

    # ---------------------------------------------------------------- stage layout
    def _vp_size(self) -> int:
        return max(1, int(self.strategy.interleaving_size))

    def _vpp_chunk_name(self, stage_name: str, virtual_rank: int) -> str:
        return f"{stage_name}_v{virtual_rank}"

    def get_num_layers_to_build(self, config: StrategyConfig, model_conf: ModelConfig, parallel_stage: str = "first",
                                virtual_pp_rank: int = None) -> int:
        """Transformer layers owned by a pipeline stage (or one of its virtual chunks).

        Handles Megatron's uneven first/last stage and the embedding / loss "count as a layer" options
        (reference perf_llm.py:539-674).
        """
        uneven = (config.num_layers_in_first_pipeline_stage is not None
                  or config.num_layers_in_last_pipeline_stage is not None)
        if uneven:
            assert not (config.account_for_embedding_in_pipeline_split or config.account_for_loss_in_pipeline_split), \
                "Does not support standalone embedding stage and standalone loss stage with uneven pp"
            todo, stages_left = model_conf.layer_num, config.pp_size
            if config.num_layers_in_first_pipeline_stage is not None:
                todo -= config.num_layers_in_first_pipeline_stage
                stages_left -= 1
            if config.num_layers_in_last_pipeline_stage is not None:
                todo -= config.num_layers_in_last_pipeline_stage
                stages_left -= 1
            if stages_left > 0:
                assert todo % stages_left == 0, (
                    "With uneven pipelineing the left over layers must be divisible by left over stages, "
                    f"layers_to_distribute={todo}, pipeline_stages_left={stages_left}")
                per_rank = todo // stages_left
            else:
                per_rank = 0
            if parallel_stage == "first" and config.num_layers_in_first_pipeline_stage is not None:
                per_rank = config.num_layers_in_first_pipeline_stage
            if parallel_stage == "last" and config.num_layers_in_last_pipeline_stage is not None:
                per_rank = config.num_layers_in_last_pipeline_stage
        else:
            total = model_conf.layer_num + int(config.account_for_embedding_in_pipeline_split) \
                + int(config.account_for_loss_in_pipeline_split)
            assert total % config.pp_size == 0, \
                f"num_layers should be divisible by pipeline_model_parallel_size, but got {total} and {config.pp_size}"
            per_rank = total // config.pp_size

        if virtual_pp_rank is None:
            n = per_rank
            if parallel_stage == "first" and config.account_for_embedding_in_pipeline_split:
                n -= 1
                assert n >= 0, "Not enough layers in the first pipeline stage"
            if parallel_stage == "last" and config.account_for_loss_in_pipeline_split:
                n -= 1
                assert n >= 0, "Not enough layers in the last pipeline stage"
            return n
        vp = max(1, int(config.interleaving_size))
        assert 0 <= virtual_pp_rank < vp, f"virtual_pp_rank should be in [0, {vp - 1}], got {virtual_pp_rank}"
        assert per_rank % vp == 0, \
            f"num_layers_per_pipeline_rank={per_rank} must be divisible by interleaving_size={vp}"
        n = per_rank // vp
        if parallel_stage == "first" and config.account_for_embedding_in_pipeline_split and virtual_pp_rank == 0:
            n -= 1
        if parallel_stage == "last" and config.account_for_loss_in_pipeline_split and virtual_pp_rank == vp - 1:
            n -= 1
        assert n >= 0, f"Not enough layers in virtual pipeline stage: stage={parallel_stage}, virtual_pp_rank={virtual_pp_rank}"
        return n

    def _new_chunk(self, layer_num, dense_layers, preprocess, postprocess, specific_name) -> StageChunk:
        return StageChunk(layer_num=layer_num, dense_layers=dense_layers, preprocess=preprocess,
                          postprocess=postprocess, model_config=self.model_config, strategy=self.strategy,
                          system=self.system, specific_name=specific_name,
                          block_cache=getattr(self, "_block_cache", None))

    def build(self):
        """Create first / middle / last stage chunks (and virtual chunks when interleaving)."""
        self.strategy.sanity_check()
        s, pp = self.strategy, self.strategy.pp_size
        self.model_chunk_dict = {}
        self.vpp_chunk_dict = {}
        self.vpp_stage_chunk_names = {FIRST_CHUNK: [], MIDDLE_CHUNK: [], LAST_CHUNK: []}
        remaining_dense = self.model_config.dense_layers
        stage_dense = {}

        d = max(0, remaining_dense)
        remaining_dense -= d
        stage_dense[FIRST_CHUNK] = d
        self.model_chunk_dict[FIRST_CHUNK] = self._new_chunk(
            self.get_num_layers_to_build(s, self.model_config, "first"), d, True, pp == 1, "GPTModel_first_pp_stage")
        if pp > 2:
            d = max(0, remaining_dense)
            remaining_dense -= d * (pp - 2)
            stage_dense[MIDDLE_CHUNK] = d
            self.model_chunk_dict[MIDDLE_CHUNK] = self._new_chunk(
                self.get_num_layers_to_build(s, self.model_config, "middle"), d, False, False, "GPTModel_middle_pp_stage")
        if pp > 1:
            d = max(0, remaining_dense)
            stage_dense[LAST_CHUNK] = d
            self.model_chunk_dict[LAST_CHUNK] = self._new_chunk(
                self.get_num_layers_to_build(s, self.model_config, "last"), d, False, True, "GPTModel_last_pp_stage")

        vp = self._vp_size()
        if vp > 1:
            flags = {FIRST_CHUNK: (True, pp == 1), MIDDLE_CHUNK: (False, False), LAST_CHUNK: (False, True)}
            for key in (FIRST_CHUNK, MIDDLE_CHUNK, LAST_CHUNK):
                if key not in self.model_chunk_dict:
                    continue
                pre, post = flags[key]
                for v in range(vp):
                    n_layers = self.get_num_layers_to_build(s, self.model_config, _STAGE_NAME[key], virtual_pp_rank=v)
                    cname = self._vpp_chunk_name(key, v)
                    self.vpp_chunk_dict[cname] = self._new_chunk(
                        n_layers, stage_dense.get(key, 0) if v == 0 else 0, pre and v == 0, post and v == vp - 1,
                        f"{cname}_model")
                    self.vpp_stage_chunk_names[key].append(cname)

    # ---------------------------------------------------------------- evaluation
    def _chunk_input_shape(self, preprocess: bool):
        s = self.strategy
        if preprocess:
            return [s.micro_batch_size, s.seq_len // s.cp_size]
        seq = s.seq_len // s.tp_size if s.enable_sequence_parallel else s.seq_len
        return [s.micro_batch_size, seq // s.cp_size, self.model_config.hidden_size]

    def _run(self):
        self.pp_state_peak_point = {}
        for name, chunk in self.model_chunk_dict.items():
            chunk(self._chunk_input_shape(chunk.preprocess))
            self.pp_state_peak_point[name] = chunk.compute_activations()
        if self._vp_size() > 1:
            for names in self.vpp_stage_chunk_names.values():
                for cname in names:
                    chunk = self.vpp_chunk_dict[cname]
                    chunk(self._chunk_input_shape(chunk.preprocess))
                    self.pp_state_peak_point[cname] = chunk.compute_activations()

    def capture(self, save_path):
        from .graph import capture_graph
        os.makedirs(save_path, exist_ok=True)
        graph = capture_graph(self)
        graph.export_json(os.path.join(save_path, "model_graph.json"))
        return graph

    @property
    def global_hidden_states_size(self):
        return self.strategy.global_batch_size * self.strategy.seq_len * self.model_config.hidden_size

    @property
    def micro_hidden_states_size(self):
        return self.strategy.micro_batch_size * self.strategy.seq_len * self.model_config.hidden_size

    # ---------------------------------------------------------------- optimizer / data-parallel time
    def _compute_optim_time(self, model_name):
        """Megatron distributed-optimizer step as a chain of memory-bound kernels (each pays launch latency):
        zero-grad, l2-norm, scale, shard norm + clip, Adam, fp32->bf16 param copy.  (reference :1470-1504)"""
        info = self.model_chunk_dict[model_name].get_model_info()
        states = info.all_state_bytes
        G = info.all_grad_bytes
        mem = lambda b: self.system.compute_mem_access_time("default", b)
        res = {"optim_time": 0, "optim_exposed_time": 0}
        if getattr(self.strategy, "fused_optimizer_model", False):
            # this repo's runtime (csrc/elementwise.cu: sqnorm_kernel + adamw_kernel): one pass over the grad shard for the
            # norm, one pass that reads grad + master/m/v and writes master/m/v + the bf16 parameter and clears the grad
            # priced from the system config's "optimizer" bandwidth entry when it has one (long fp32 streams run closer to the copy
            # bandwidth than the short elementwise kernels behind "default"; calib.measure.measure_bandwidth_ops)
            mem = lambda b: self.system.compute_mem_access_time("optimizer", b)
            shard_grads = states / 6 if self.strategy.grad_reduce_in_bf16 else states / 3
            res["l2_norm_after_reduce_time"] = mem(shard_grads)
            res["adam_time"] = mem(shard_grads + states + states + (states / 3) * 0.5 + shard_grads)
            if self.strategy.dp_size * self.strategy.cp_size > 1:
                res["zero_grad_buffer_time"] = mem(G)            # the full main-grad buffer, after its reduce-scatter
            total = sum(res.values())
            res["optim_time"] = res["optim_exposed_time"] = total
            return res
        res["zero_grad_buffer_time"] = mem(G)
        res["l2_norm_before_reduce_time"] = mem(G)
        res["mul_before_reduce_time"] = mem(2 * G) if self.strategy.dp_size * self.strategy.cp_size > 1 else 0
        grads_chunk = states / 6 if self.strategy.grad_reduce_in_bf16 else states / 3
        weight_bytes = states / 3
        res["l2_norm_after_reduce_time"] = mem(grads_chunk)
        res["grads_clip_after_reduce_time"] = mem(2 * grads_chunk)
        res["adam_time"] = mem(grads_chunk + 3 * states)
        res["copy_main_params_to_model_params_time"] = mem(weight_bytes + 0.5 * weight_bytes)
        total = sum(res.values())
        res["optim_time"] = total
        res["optim_exposed_time"] = total
        return res

    def _compute_dp_time(self, model_name):
        """ZeRO-1: bucketed reduce-scatter of main grads + all-gather of updated params (ZeRO-0: all-reduce).
        bucket = max(40e6, 1e6 * group) * 4 bytes; dense group = dp*cp, expert group = edp.  (reference :1513-1597)"""
        chunk = self.model_chunk_dict[model_name]
        s = self.strategy
        is_moe = self.model_config.model_type == "moe"

        def to_param_bytes(grad_bytes):
            return grad_bytes / chunk.main_grad_element_size * self.dtype_to_element_size[s.dtype]

        def family(rs_bytes, ag_bytes, net, group_size, tag):
            bucket = max(40000000, 1000000 * group_size) * 4
            n_rs = (rs_bytes - 1) // bucket + 1
            n_ag = (ag_bytes - 1) // bucket + 1
            if is_moe:
                n_ag *= 2
            res = {"dp_comm_time": 0, "dp_comm_exposed_time": 0}
            total, details = 0, {}
            if s.zero_state >= 1:
                rs = n_rs * self.system.compute_net_op_time("reduce_scatter", bucket, comm_num=group_size, net=net,
                                                            comm_stage=tag, strategy=s)
                ag = n_ag * self.system.compute_net_op_time("all_gather", bucket, comm_num=group_size, net=net,
                                                            comm_stage=tag, strategy=s)
                total += ag + rs
                details = {"reduce_scatter_time": rs, "all_gather_time": ag}
            else:
                total += n_rs * self.system.compute_net_op_time("all_reduce", bucket, comm_num=group_size, net=net,
                                                                comm_stage=tag, strategy=s)
            res["dp_comm_rs_size"] = rs_bytes if group_size > 1 else 0
            res["dp_comm_ag_size"] = ag_bytes if group_size > 1 else 0
            res["dp_comm_num_gather"] = 2 if is_moe else 1
            res["dp_comm_time"] = total
            res["dp_comm_exposed_time"] = total
            if details and getattr(s, "dp_overlap_model", False) and group_size > 1 and n_rs > 0 and n_ag > 0:
                # overlap extension: each collective is exposed only past the compute it can hide behind, and never
                # less than its last bucket (which has nothing left to overlap with)
                fwd_t, bwd_t = self._compute_single_batch_fwd_bwd_time(model_name)
                eta = float(getattr(s, "dp_overlap_efficiency", 0.5))    # overlapped kernels share SMs / HBM with the GEMMs
                rs_t, ag_t = details["reduce_scatter_time"], details["all_gather_time"]
                exposed = max(rs_t - eta * min(rs_t, bwd_t), rs_t / n_rs) + max(ag_t - eta * min(ag_t, fwd_t), ag_t / n_ag)
                res["dp_comm_exposed_time"] = min(total, exposed)
            if details:
                res["details"] = details
            return res

        info = chunk.get_model_info()
        dense = family(info.dense_grad_bytes, to_param_bytes(info.dense_grad_bytes), s.dp_net,
                       s.dp_size * s.cp_size, "dp_cp")
        moe = family(info.moe_grad_bytes, to_param_bytes(info.moe_grad_bytes), s.edp_net, s.edp_size, "edp")
        return {"dp_comm_exposed_time": dense["dp_comm_exposed_time"] + moe["dp_comm_exposed_time"],
                "dense": dense, "moe": moe}

    # ---------------------------------------------------------------- memory
    def _analysis_mem_impl(self, micro_batch_num, model_name=FIRST_CHUNK):
        """peak = static model memory + (k-1) cached micro-batches + peak live activation of one micro-batch."""
        info = self.model_chunk_dict[model_name].get_model_info()
        s = self.strategy
        result = {
            "micro_batch_num": s.micro_batch_num, "micro_batch_size": s.micro_batch_size,
            "cached_micro_batch_num": micro_batch_num - 1,
            "parallel_config": {
                "parallelism": s.parallelism, "fp8": s.fp8,
                "recompute_status": {
                    "layer_num": self.model_config.layer_num,
                    "actual_layer_num": self.model_chunk_dict[FIRST_CHUNK].layer_num,
                    "recompute_layer": s.recompute_layer_num,
                    "recompute_recompute_granularity": s.recompute_granularity}},
        }
        dense = dict(all_mem=info.dense_weight_bytes + info.dense_grad_bytes + info.dense_state_bytes,
                     detail=dict(weight_bytes=info.dense_weight_bytes, grad_bytes=info.dense_grad_bytes,
                                 state_bytes=info.dense_state_bytes))
        moe = dict(all_mem=info.moe_weight_bytes + info.moe_grad_bytes + info.moe_state_bytes,
                   detail=dict(weight_bytes=info.moe_weight_bytes, grad_bytes=info.moe_grad_bytes,
                               state_bytes=info.moe_state_bytes))
        dummy = dict(all_mem=info.te_dummy_wgrad_bytes,
                     detail=dict(dummy_wgrad_bytes=info.te_dummy_wgrad_bytes,
                                 shape_count=len(info.te_dummy_wgrad_shapes),
                                 shapes=[list(x) for x in sorted(info.te_dummy_wgrad_shapes)]))
        model_mem = dense["all_mem"] + moe["all_mem"] + dummy["all_mem"]
        result["model_mem"] = model_mem
        result["model_mem_detail"] = dict(dense=dense, moe=moe, te_dummy_wgrad=dummy)
        pk: PeakPoint = self.pp_state_peak_point[model_name]
        result["fwd_activation_cache_per_micro_batch"] = f"{pk.activation_mem_cache / 1024 / 1024 / 1024:.4f} GB"
        result["peak_activation_mem_in_1F1B"] = pk.peak_mem
        result["peak_mem"] = model_mem + (micro_batch_num - 1) * pk.activation_mem_cache + pk.peak_mem
        result["peak_mem_with_reserved"] = result["peak_mem"] / s.mem_factor
        result["memory_reserved_ratio"] = str(s.mem_factor)
        result["peak_path"] = f"{pk.peak_path}, stage=[{pk.peak_stage}]"
        result["raw"] = None  # placeholder so raw numbers survive the human formatting below
        raw = {"model_mem": model_mem, "peak_mem": result["peak_mem"],
               "peak_mem_with_reserved": result["peak_mem_with_reserved"],
               "peak_activation_mem_in_1F1B": pk.peak_mem, "activation_cache_per_micro_batch": pk.activation_mem_cache}
        convert_final_result_to_human_format(result)
        result["raw"] = raw
        return result

    def analysis_mem(self) -> Result:
        pp = self.strategy.pp_size
        if self._vp_size() > 1 and self.vpp_stage_chunk_names.get(FIRST_CHUNK) and not self.strategy.pp_comm_async:
            from .schedule_vpp import analysis_sync_vpp_stage_mem
            if pp == 1:
                return Result(analysis_sync_vpp_stage_mem(self, 0))
            return Result({self._vpp_stage_result_key(r): analysis_sync_vpp_stage_mem(self, r) for r in range(pp)})
        if pp == 1:
            return Result(self._analysis_mem_impl(1, FIRST_CHUNK))
        result = {"first_stage": self._analysis_mem_impl(pp, FIRST_CHUNK)}
        if pp > 2:
            result["middle_stage"] = self._analysis_mem_impl(pp - 1, MIDDLE_CHUNK)
        result["last_stage"] = self._analysis_mem_impl(1, LAST_CHUNK)
        return Result(result)

    def _stage_key_for_pp_rank(self, pp_rank: int):
        if pp_rank == 0:
            return FIRST_CHUNK
        if pp_rank == self.strategy.pp_size - 1:
            return LAST_CHUNK
        return MIDDLE_CHUNK

    def _vpp_stage_result_key(self, pp_rank: int):
        if self.strategy.pp_size <= 1 or pp_rank == 0:
            return "first_stage"
        if pp_rank == self.strategy.pp_size - 1:
            return "last_stage"
        return f"pp_stage_{pp_rank}"            # reference perf_llm.py:1688-1693

    def get_pp_stage_peak_mem(self, mem_result, peak_mem_key, toG: bool = False):
        assert peak_mem_key in ("peak_mem_with_reserved", "peak_mem"), \
            f"peak_mem_key should be in ['peak_mem_with_reserved', 'peak_mem'] but got {peak_mem_key}"
        data = mem_result.data if isinstance(mem_result, Result) else mem_result

        def parse(v):
            b = HumanReadableSize.from_string(v, HumanReadableSize.BYTE_UNITS, 1024, target_unit="B").get_value()
            return b / 1024 ** 3 if toG else b
        if "peak_mem" in data:
            return dict(first_stage=parse(data.get(peak_mem_key)))
        return {k: parse(v.get(peak_mem_key)) for k, v in data.items() if isinstance(v, dict) and peak_mem_key in v}

    # ---------------------------------------------------------------- time
    def _chunk_stage_key(self, model_name):
        if model_name in (FIRST_CHUNK, MIDDLE_CHUNK, LAST_CHUNK):
            return model_name
        for key, names in self.vpp_stage_chunk_names.items():
            if model_name in names:
                return key
        return model_name

    def _compute_single_batch_phase_inputs(self, model_name):
        """Per-micro-batch phase durations of a stage: recv / compute(+all intra-layer comm, exposed) / send."""
        chunk = self.model_chunk_dict.get(model_name) or self.vpp_chunk_dict.get(model_name)
        if chunk is None:
            raise KeyError(f"Unknown model chunk: {model_name}")
        c = chunk.get_cost_info()
        s = self.strategy
        p2p = 0.0
        if s.pp_size > 1:
            size = get_pp_p2p_comm_size(s, self.model_config.hidden_size, self.dtype_to_element_size[s.dtype])
            p2p = self.system.compute_net_op_time("p2p", size, 2, net=s.pp_net)
        key = self._chunk_stage_key(model_name)
        if s.pp_size <= 1:
            fr = fs = br = bs = 0.0
        elif key == FIRST_CHUNK:
            fr, fs, br, bs = 0.0, p2p, p2p, 0.0
        elif key == LAST_CHUNK:
            fr, fs, br, bs = p2p, 0.0, 0.0, p2p
        else:
            fr = fs = br = bs = p2p
        return {"fwd_recv": fr, "fwd_compute": c.fwd_compute_time + c.fwd_net_time, "fwd_send": fs, "bwd_recv": br,
                "bwd_compute": c.bwd_compute_time + c.bwd_net_time + c.recompute_compute_time + c.recompute_net_time,
                "bwd_send": bs}

    def _compute_single_batch_fwd_bwd_time(self, model_name, chunk=False):
        ph = self._compute_single_batch_phase_inputs(model_name)
        f = ph["fwd_recv"] + ph["fwd_compute"] + ph["fwd_send"]
        b = ph["bwd_recv"] + ph["bwd_compute"] + ph["bwd_send"]
        return f + b if chunk else (f, b)

    def _stage_phase_list(self):
        pp = self.strategy.pp_size
        phases = [self._compute_single_batch_phase_inputs(FIRST_CHUNK)]
        if pp > 2:
            phases.extend([self._compute_single_batch_phase_inputs(MIDDLE_CHUNK)] * (pp - 2))
        if pp > 1:
            phases.append(self._compute_single_batch_phase_inputs(LAST_CHUNK))
        return phases

    def _compute_pp_total_time(self, return_schedules=False):
        if self._vp_size() > 1 and self.vpp_stage_chunk_names.get(FIRST_CHUNK):
            if self.strategy.pp_comm_async:
                if getattr(self.strategy, "vpp_async_perf_via_simulation", False):
                    from ..sim.runner import simulate_schedule       # extension: time the schedule with the event engine
                    return simulate_schedule(self, return_schedules=return_schedules)
                raise RuntimeError("analysis_cost() does not support async VPP perf timing yet; "
                                   "use pp_comm_async=False for perf timing or simulate() for async VPP traces")
            from .schedule_vpp import compute_interleaved_sync_schedule
            return compute_interleaved_sync_schedule(self, return_schedules=return_schedules)
        return one_f_one_b(self.strategy.pp_size, self.strategy.micro_batch_num, self._stage_phase_list(),
                           return_schedules=return_schedules)

    def calculate_1f1b_bubble(self, pp, mbc, forward_times, backward_times, draw=False, stage_phases=None,
                              return_schedules=False):
        from .schedule import one_f_one_b_simple
        if stage_phases is None:
            out = one_f_one_b_simple(pp, mbc, forward_times, backward_times, return_schedules=True)
        else:
            out = one_f_one_b(pp, mbc, stage_phases, return_schedules=True)
        total, sched = out
        if draw:
            from .trace_export import export_pipeline_schedule_trace
            path = draw if isinstance(draw, str) else os.path.abspath("corrected_1F1B_pipeline_trace.json")
            export_pipeline_schedule_trace(sched, path, title=f"perf_1f1b_pp{pp}_mbc{mbc}")
        return (total, sched) if return_schedules else total

    def export_pp_schedule_trace(self, output_json_path):
        """Chrome trace of the analytic pipeline schedule (``traceEvents`` + displayTimeUnit ms)."""
        from .trace_export import export_pipeline_schedule_trace
        total, sched = self._compute_pp_total_time(return_schedules=True)
        s = self.strategy
        export_pipeline_schedule_trace(sched, output_json_path,
                                       title=f"perf_pp{s.pp_size}_vp{self._vp_size()}_mbc{s.micro_batch_num}")
        return total

    def _single_batch_stats(self, model_name):
        chunk = self.model_chunk_dict[model_name]
        ci, co = chunk.get_compute_info(), chunk.get_cost_info()
        rc = self.strategy.enable_recompute
        cost = {"fwd_compute_time": co.fwd_compute_time, "bwd_compute_time": co.bwd_compute_time,
                "recompute_compute_time": co.recompute_compute_time, "fwd_net_time": co.fwd_net_time,
                "bwd_net_time": co.bwd_net_time, "recompute_net_time": co.recompute_net_time,
                "fwd_net_exposed_time": co.fwd_net_exposed_time, "bwd_net_exposed_time": co.bwd_net_exposed_time,
                "recompute_net_exposed_time": co.recompute_net_exposed_time, "fwd_time": co.fwd_time,
                "bwd_time": co.bwd_time, "recompute_time": co.recompute_time if rc else 0}
        comp = {"fwd_flops": ci.fwd_flops, "recompute_flops": ci.recompute_flops if rc else 0,
                "bwd_flops": ci.bwd_flops, "fwd_accessed_mem": ci.fwd_accessed_mem,
                "recompute_accessed_mem": ci.recompute_accessed_mem if rc else 0,
                "bwd_accessed_mem": ci.bwd_accessed_mem}
        return {"compute_info": comp, "cost_info": cost}

    def _gbs_compute(self, stat, model_name):
        n = self.strategy.micro_batch_num
        out = {"batch_compute_stat": stat,
               "fwd_compute_time": stat["cost_info"]["fwd_compute_time"] * n,
               "recompute_time": stat["cost_info"]["recompute_compute_time"] * n,
               "bwd_compute_time": stat["cost_info"]["bwd_compute_time"] * n,
               "optim_time": self._compute_optim_time(model_name),
               "fwd_flops": stat["compute_info"]["fwd_flops"] * n,
               "recompute_flops": stat["compute_info"]["recompute_flops"] * n,
               "bwd_flops": stat["compute_info"]["bwd_flops"] * n}
        out["model_flops"] = out["fwd_flops"] + out["bwd_flops"]
        return out

    def _gbs_comm(self, stat, model_name):
        n = self.strategy.micro_batch_num
        dp = self._compute_dp_time(model_name)
        intra = sum(stat["cost_info"][k] for k in ("fwd_net_time", "bwd_net_time", "recompute_net_time"))
        inter = 0
        if self.strategy.pp_size > 1:
            ph = self._compute_single_batch_phase_inputs(model_name)
            inter = ph["fwd_recv"] + ph["fwd_send"] + ph["bwd_recv"] + ph["bwd_send"]
        return {"dp_comm_time": dp,
                "intra_comm_time": {"intra_exposed_time_per_batch": intra, "intra_exposed_time": intra * n},
                "inter_comm_time": {"inter_exposed_time_per_batch": inter, "inter_exposed_time": inter * n}}

    @staticmethod
    def _breakdown(comp, comm):
        return {"fwd_compute_time": comp["fwd_compute_time"], "recompute_time": comp["recompute_time"],
                "bwd_compute_time": comp["bwd_compute_time"], "optim_time": comp["optim_time"]["optim_exposed_time"],
                "intra_exposed_time": comm["intra_comm_time"]["intra_exposed_time"],
                "inter_exposed_time": comm["inter_comm_time"]["inter_exposed_time"],
                "dp_exposed_time": comm["dp_comm_time"]["dp_comm_exposed_time"]}

    def _analysis_single_iter_cost_impl(self):
        """iteration = straggler * pipeline_schedule_time + max over stages of (DP comm + optimizer)."""
        s, pp = self.strategy, self.strategy.pp_size
        res = {}
        stat = self._single_batch_stats(FIRST_CHUNK)
        comm_first = self._gbs_comm(stat, FIRST_CHUNK)
        comp_first = self._gbs_compute(stat, FIRST_CHUNK)
        chunk_time = self._compute_single_batch_fwd_bwd_time(FIRST_CHUNK, chunk=True)
        chunk_time_last = chunk_time_mid = 0
        if pp > 1:
            stat_l = self._single_batch_stats(LAST_CHUNK)
            comm_last = self._gbs_comm(stat_l, LAST_CHUNK)
            comp_last = self._gbs_compute(stat_l, LAST_CHUNK)
            chunk_time_last = self._compute_single_batch_fwd_bwd_time(LAST_CHUNK, chunk=True)
            if pp > 2:
                chunk_time_mid = self._compute_single_batch_fwd_bwd_time(MIDDLE_CHUNK, chunk=True)
            res["breakdown_result_last_stage"] = self._breakdown(comp_last, comm_last)

        def numel(attr):
            n = getattr(self.model_chunk_dict[FIRST_CHUNK]._model_info, attr)
            n += (getattr(self.model_chunk_dict[MIDDLE_CHUNK]._model_info, attr) if pp > 2 else 0) * (pp - 2)
            n += getattr(self.model_chunk_dict[LAST_CHUNK]._model_info, attr) if pp > 1 else 0
            return n
        dense_numel, moe_numel = numel("weight_numel"), numel("moe_weight_numel")

        def dp_and_optim(name):
            return self._compute_dp_time(name)["dp_comm_exposed_time"] + self._compute_optim_time(name)["optim_exposed_time"]

        sched_time = self._compute_pp_total_time()
        if s.enable_straggler_model:
            ratio = straggler_increase_ratio(effective_straggler_samples(
                s.world_size, self.system.num_per_node, s.dp_size, s.edp_size))
        else:
            ratio = 1.0
        sched_time = ratio * sched_time
        durations = [sched_time + dp_and_optim(FIRST_CHUNK)]
        if pp > 2:
            durations.append(sched_time + dp_and_optim(MIDDLE_CHUNK))
        if pp > 1:
            durations.append(sched_time + dp_and_optim(LAST_CHUNK))
        iter_ms = max(durations)
        tokens = s.seq_len * s.global_batch_size
        fpt = self.model_config.flops_per_token(context_seq_len=s.seq_len, with_attn=True)
        theory_flops = fpt * tokens // s.world_size
        tgs = tokens / (iter_ms / 1000) / s.world_size
        tflops = theory_flops / (iter_ms / 1000) / 1e12
        tflops_per_token = fpt / (iter_ms / 1000) / 1e12
        mfu = tflops / self.system.accelerator.op["default"].tflops
        mbc = s.micro_batch_num

        res["comm_details"] = comm_first
        res["compute_details"] = comp_first
        res["breakdown_result"] = self._breakdown(comp_first, comm_first)
        res["all_tokens_per_iter"] = tokens
        res["straggle_ratio"] = ratio

        def chunk_entry(name, ctime, dur):
            return {name: {"duration_time(chunk_timexmbc+dp_optim+bubble)": dur, "chunk_time(fwd+bwd)": ctime,
                           "dp_and_optim_time": dp_and_optim(name),
                           "bubble_time": sched_time / ratio - mbc * ctime,
                           "straggler_time": sched_time * (ratio - 1) / ratio}}
        res["all_chunk_times"] = chunk_entry(FIRST_CHUNK, chunk_time, durations[0])
        if pp > 2:
            res["all_chunk_times"].update(chunk_entry(MIDDLE_CHUNK, chunk_time_mid, durations[1]))
        if pp > 1:
            res["all_chunk_times"].update(chunk_entry(LAST_CHUNK, chunk_time_last, durations[-1]))
        res.update({"duration_time_per_iter": iter_ms, "throughput_per_accelerator": tgs,
                    "throughput per GPU (TFLOP/s/GPU)": tflops,
                    "throughput per GPU per token (TFLOP/s/GPU/token)": tflops_per_token,
                    "mfu_6nd_with_attn": mfu, "mfu": mfu, "moe_param_numel": f"{moe_numel / 1e9:.2f}B"})
        res["flops_info"] = {"theory_flops": theory_flops, "model_flops": comp_first["model_flops"]}
        res["param_numel_info"] = {"dense": f"{dense_numel / 1e9:.2f}B", "moe": f"{moe_numel / 1e9:.2f}B",
                                   "all": f"{(dense_numel + moe_numel) / 1e9:.2f}B"}
        if self.model_config.model_type == "moe":
            act = dense_numel + moe_numel * (self.model_config.topk / self.model_config.expert_num)
            res["param_numel_info"].update({"activations": f"{act / 1e9:.2f}B",
                                            "activations_ratio": f"{act / (dense_numel + moe_numel) * 100:.2f}%"})
        else:
            res["param_numel_info"].update({"activations": res["param_numel_info"]["all"],
                                            "activations_ratio": f"{100:.2f}%"})
        raw = {"duration_time_per_iter_ms": iter_ms, "schedule_time_ms": sched_time,
               "tokens_per_iter": tokens, "tokens_per_sec": tokens / (iter_ms / 1000),
               "theory_flops": theory_flops}
        convert_final_result_to_human_format(res)
        res["raw"] = raw
        return res

    def analysis_cost(self) -> Result:
        return Result(self._analysis_single_iter_cost_impl())

    def analysis_gemm_costs(self):
        def merge(a, b):
            for k in a:
                a[k].extend(b[k])
            return a
        costs = self.model_chunk_dict[FIRST_CHUNK].get_all_gemm_cost_info()
        if self.strategy.pp_size > 1:
            costs = merge(costs, self.model_chunk_dict[LAST_CHUNK].get_all_gemm_cost_info())
        if self.strategy.pp_size > 2:
            mid = self.model_chunk_dict[MIDDLE_CHUNK].get_all_gemm_cost_info()
            for _ in range(self.strategy.pp_size - 2):
                costs = merge(costs, mid)
        return costs

    def analysis_op_info(self):
        return {k: c.get_all_gemm_cost_info() for k, c in self.model_chunk_dict.items()}

    # ---------------------------------------------------------------- results
    def _write_results(self, save_path, mem_result: Result, compute_result: Result, with_strategy=True):
        os.makedirs(save_path, exist_ok=True)
        arch = str(dict(self.model_chunk_dict))            # {'first_stage_chunk': LLMModel(...), ...} like the reference
        base_info = {"arch": arch, "all_param": self.model_config.param_numel,
                     "act_param": self.model_config.activated_param_numel}
        dump = lambda o: json.dumps(o, indent=2, sort_keys=False, ensure_ascii=False, default=str)
        files = {"model_arch": arch, "base_info.json": dump(base_info), "mem_result.json": dump(mem_result.data),
                 "compute_result.json": dump(compute_result.data), "system_config.json": str(self.system),
                 "net_info.json": json.dumps(self.system.real_comm_bw, indent=4),
                 "model_config.json": str(self.model_config)}
        if with_strategy:
            files["strategy_config.json"] = str(self.strategy)
        for fname, text in files.items():
            with open(os.path.join(save_path, fname), "w", encoding="utf-8") as fh:
                fh.write(text)

    def _dump_memory_and_cost(self, mem_result, compute_result, save_path):
        mem_result = mem_result if isinstance(mem_result, Result) else Result(mem_result)
        compute_result = compute_result if isinstance(compute_result, Result) else Result(compute_result)
        self._write_results(save_path, mem_result, compute_result, with_strategy=False)

    def analysis(self, save_path=None, console_log=True):
        """Memory + cost analysis, optional result files, console summary; returns the headline dict."""
        mem_result = self.analysis_mem()
        compute_result = self.analysis_cost()
        if SIMU_CHECK:
            save_path = TMP_PATH
        if save_path is not None:
            print(f"Saving analysis results to {save_path}")
            self._write_results(save_path, mem_result, compute_result)
        md, cd = mem_result.data, compute_result.data
        if "peak_mem" in md:
            peak_mem, peak_res = md["peak_mem"], md["peak_mem_with_reserved"]
        else:
            peak_mem = {k: v["peak_mem"] for k, v in md.items()}
            peak_res = {k: v["peak_mem_with_reserved"] for k, v in md.items()}
        s, m = self.strategy, self.model_config
        if console_log:
            act = f", act={cd['param_numel_info']['activations']}" if m.model_type == "moe" else ""
            print(f"-------------SIMUMAX SUMMARY  \033[33m{m.model_name}({cd['param_numel_info']['all']}{act}) "
                  f"TP={s.tp_size},EP={s.ep_size},PP={s.pp_size}\033[0m -------------")
            print(f"- parallelism = layer{m.layer_num}.dense{m.dense_layers}.{s.parallelism}")
            print(f"- recompute = {s.recompute_status}")
            print(f"- \033[31mdtype = {'fp8' if s.fp8 else 'bf16'}, grad_reduce = {'bf16' if s.grad_reduce_in_bf16 else 'fp32'}\033[0m")
            print(f"- system = {self.system.sys_name}")
            print(f"- model_type = {m.model_type}")
            print(f"· \033[32mmfu = {cd['mfu_6nd_with_attn']:.2f}\033[0m")
            print(f"· \033[32mTFLOPS = {cd['throughput per GPU (TFLOP/s/GPU)']:.2f}T "
                  f"(tflops={cd['flops_info']['theory_flops']}, duration={cd['duration_time_per_iter']})\033[0m")
            print(f"· \033[32mTFLOPS_PER_TOKEN = {cd['throughput per GPU per token (TFLOP/s/GPU/token)']:.2f}T, "
                  f"duration={cd['duration_time_per_iter']})\033[0m")
            print(f"· \033[31mpeak_alloc_mem = {peak_mem}\033[0m")
            print(f"- peak_alloc_mem_with_reserved = {peak_res}")
            print(f"- TGS_per_gpu = {cd['throughput_per_accelerator']}")
            print(f"- net = {s.net} ")
            print("------------------------------------------")
        if ENABLE_SIMU_GRAPH and save_path is not None:
            from .graph import visualize_with_graphviz
            self.capture(save_path)
            visualize_with_graphviz(os.path.join(save_path, "model_graph.json"),
                                    output_path=os.path.join(save_path, "computational_graph"))
        return {"model": m.model_name, "model_type": m.model_type, "params": cd["param_numel_info"]["all"],
                "system": self.system.sys_name, "peak_mem": peak_mem, "peak_mem_with_reserved": peak_res,
                "duration_time_per_iter": cd["duration_time_per_iter"],
                "TFLOPS": cd["throughput per GPU (TFLOP/s/GPU)"], "TGS_per_gpu": cd["throughput_per_accelerator"],
                "mfu": cd["mfu_6nd_with_attn"], "parallelism": s.parallelism, "recompute": s.recompute_status,
                "dtype": f"{'fp8' if s.fp8 else 'bf16'},grad_reduce in {'bf16' if s.grad_reduce_in_bf16 else 'fp32'}",
                "net": s.net}

    def perf(self, save_path=None, console_log=True):
        """Convenience: ``run_estimate()`` + ``analysis()`` in one call (the team's word for it)."""
        self.run_estimate()
        return self.analysis(save_path=save_path, console_log=console_log)

    def estimate_dualpipe(self, model_name=None) -> dict:
        """DualPipe / ZB1P what-if for the configured job (extension: the reference ships the closed forms in pp_simu/utils.py
        but does not connect them to the model).  Takes one stage chunk's evaluated ops: F = forward compute + collectives,
        B = full backward, W = its weight-gradient part, and the mutually overlapped F&B pair layer by layer (compute of one
        direction hides the expert / tensor-parallel collectives of the other); returns bubbles, durations and MFUs of 1F1B,
        ZB1P and DualPipe for the strategy's ``pp_size`` and ``micro_batch_num`` (schedule time only, no optimizer)."""
        from ..pp_simu.utils import cal_FandB_multi_layers, perf_dualpp
        if model_name is None:       # the slowest stage paces a pipeline; the closed forms assume identical stages
            model_name = max(self.model_chunk_dict, key=lambda k: self._compute_single_batch_fwd_bwd_time(k, chunk=True))
        chunk = self.model_chunk_dict[model_name]
        s = self.strategy
        per_layer = {}
        f = b = w = 0.0
        for n in chunk.nodes:
            c = n.cost
            nf = c.fwd_compute_time
            nb = c.bwd_grad_act_time + c.bwd_grad_w_time + c.recompute_compute_time
            cf, cb = c.fwd_net_time, c.bwd_grad_act_net_time + c.bwd_grad_w_net_time + c.recompute_net_time
            f += nf + cf
            b += nb + cb
            w += c.bwd_grad_w_time
            acc = per_layer.setdefault(n.layer_idx if isinstance(n.layer_idx, int) and n.layer_idx >= 0 else -1, [0.0, 0.0, 0.0, 0.0])
            acc[0] += nf; acc[1] += nb; acc[2] += cf; acc[3] += cb
        layers = [v for _, v in sorted(per_layer.items())]
        fb = cal_FandB_multi_layers([v[0] for v in layers], [v[1] for v in layers], [v[2] for v in layers], [v[3] for v in layers])
        flops = chunk.get_compute_info().fwd_flops + chunk.get_compute_info().bwd_flops
        peak = self.system.accelerator.op["matmul"].tflops
        pp = s.pp_size if s.pp_size % 2 == 0 else s.pp_size + 1
        out = perf_dualpp(max(pp, 2), s.micro_batch_num, f, b, w, f_and_b=fb["overlapped"], flops_per_microbatch=flops, peak_tflops=peak)
        out.update({"chunk": model_name, "pp_size_used": max(pp, 2), "f_ms": f, "b_ms": b, "w_ms": w, "f_and_b_overlapped_ms": fb["overlapped"],
                    "f_and_b_serial_ms": fb["serial"]})
        return out

    def simulate(self, save_path, merge_lanes=True, memory_timeline="auto", jitter: float = 0.0, slow_ranks=None, seed: int = 0):
        """Discrete-event replay of the iteration: Chrome trace + memory timeline artifacts.

        ``memory_timeline``: "auto" follows the reference (memory artifacts only without pipeline parallelism or with blocking
        p2p, core/simu_artifacts.py:9-10); ``True`` also tracks memory under asynchronous pipeline communication (extension:
        this engine's tracker hangs off the scope events, which do not depend on the p2p mode), ``False`` switches it off.
        ``jitter`` / ``slow_ranks`` / ``seed``: fault injection (extension) -- mean-one log-normal noise on every compute atom and
        constant per-rank slow-down factors, to see how stragglers propagate through the pipeline and the collectives."""
        from ..sim.runner import run_simulation
        return run_simulation(self, save_path, merge_lanes=merge_lanes, memory_timeline=memory_timeline, jitter=jitter,
                              slow_ranks=slow_ranks, seed=seed)

    # search API lives in core/search.py and is mixed in below
from .search import SearchMixin  # noqa: E402

for _name in dir(SearchMixin):
    if not _name.startswith("__"):
        setattr(PerfLLM, _name, getattr(SearchMixin, _name))
