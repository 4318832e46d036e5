"""perf() of the deepseekv2_layer4_ep4_pp2_full_recompute case (see _cases.py for the model / strategy / system triple)."""
from _cases import run

if __name__ == "__main__":
    run("deepseekv2_layer4_ep4_pp2_full_recompute")
