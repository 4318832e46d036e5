"""perf() of the llama3_70b_layer12_tp8 case (see _cases.py for the model / strategy / system triple)."""
from _cases import run

if __name__ == "__main__":
    run("llama3_70b_layer12_tp8")
