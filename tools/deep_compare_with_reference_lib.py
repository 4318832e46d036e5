"""(library part of deep_compare_with_reference.py) Deep differential check: the COMPLETE analysis_mem() / analysis_cost() result dictionaries of the reference and of this
repository (every key, every formatted value, the peak-path strings) and the model_arch text (nested module view with
per-module time / FLOP / memory totals) for hand-picked and random configurations.
    python tools/deep_compare_with_reference.py [--cases 0,2,20] [--seed N] [--random K] [--fp8]
"""
import sys, json, io, os, random, warnings
from contextlib import redirect_stdout
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import compare_with_reference as cw, fuzz_vs_reference as fz
CFG=cw.CFG; REF=cw.REF
def run(mod_prefix, case, system):
    strategy, model, layers, mover, sover = case[:5]
    if mod_prefix == "ref":
        sys.path.insert(0, REF)
        from simumax.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax.core.perf_llm import PerfLLM
    else:
        from simumax_b200.core.config import ModelConfig, StrategyConfig, SystemConfig
        from simumax_b200.core.perf_llm import PerfLLM
    st = StrategyConfig.init_from_config_file(f"{CFG}/strategy/{strategy}.json")
    for k, v in sover.items(): setattr(st, k, v)
    p = PerfLLM(); p._search_verbose = False
    with redirect_stdout(io.StringIO()):
        p.configure(strategy_config=st, model_config=ModelConfig.init_from_config_file(f"{CFG}/models/{model}.json"),
                    system_config=SystemConfig.init_from_config_file(f"{CFG}/system/{system}.json"))
        if layers: p.model_config.layer_num = layers
        for k, v in mover.items(): setattr(p.model_config, k, v)
        p.model_config.padded_vocab_size = True; p.model_config.make_vocab_size_divisible_by = 128
        p.run_estimate()
        mem = p.analysis_mem().data; cost = p.analysis_cost().data
        arch = str(dict(p.model_chunk_dict))          # text of model_arch / base_info.json["arch"]: the nested module view
        gemm = p.analysis_op_info()                    # per-chunk GEMM tables (fwd / dgrad / wgrad rows of every linear)
        ops = {}
        for name, chunk in p.model_chunk_dict.items():
            try:
                ops[name] = chunk.analysis_op_info(True)      # per-op table with roofline details
            except AttributeError:                            # the reference cannot build it for MoE chunks (Permutation has no details)
                ops = None
                break
    out = {"mem": mem, "cost": cost, "arch": arch.split("\n"), "gemm": gemm}
    if ops is not None:
        out["ops"] = ops
    return json.loads(json.dumps(out, default=str))
SOFT=[]
def diff(a,b,path="",out=None):
    out = [] if out is None else out
    if isinstance(a,dict) and isinstance(b,dict):
        if path=="" and ("ops" in a) != ("ops" in b):            # compared only when both implementations can build it
            a={k:v for k,v in a.items() if k!="ops"}; b={k:v for k,v in b.items() if k!="ops"}
        for k in sorted(set(a)|set(b)):
            if k not in a: out.append((path+"/"+k,"<missing in ref>",str(b[k])[:60]))
            elif k not in b: out.append((path+"/"+k,str(a[k])[:60],"<missing in new>"))
            else: diff(a[k],b[k],path+"/"+k,out)
    elif isinstance(a,list) and isinstance(b,list) and len(a)==len(b):
        for i,(x,y) in enumerate(zip(a,b)): diff(x,y,f"{path}[{i}]",out)
    else:
        if a!=b:
            try:
                fa,fb=float(str(a).split()[0]),float(str(b).split()[0])
                if abs(fa-fb)<=1e-9*max(abs(fa),abs(fb),1e-300):
                    SOFT.append((path,str(a)[:60],str(b)[:60])); return out       # same number, different text (e.g. -0.0000)
            except Exception: pass
            out.append((path,str(a)[:60],str(b)[:60]))
    return out
