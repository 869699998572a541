"""Compare the traces bench.py --trace-out wrote at different GPU counts: the sharded solve may differ from the
single-GPU one by summation order only, i.e. within the step-1 bar (1e-9) while the iteration is
stable.  Prints, per pair, the worst relative cost deviation in step 1 / step 2, the first trial whose decision or
term count differs, and the final-cost deviation.

    python tools/compare_traces.py gpurun_out/trace_c5_n1.json gpurun_out/trace_c5_n2.json ...
"""
import json
import sys


def main():
    traces = [json.load(open(p)) for p in sys.argv[1:]]
    base = traces[0]
    out = []
    for t in traces[1:]:
        n = min(len(base["cost"]), len(t["cost"]))
        worst = {1: 0.0, 2: 0.0}
        first_flip = None
        for i in range(n):
            d = abs(base["cost"][i] - t["cost"][i]) / abs(base["cost"][i])
            worst[base["step"][i]] = max(worst[base["step"][i]], d)
            if first_flip is None and (base["step_is_successful"][i] != t["step_is_successful"][i] or
                                       base["linear_solver_iterations"][i] != t["linear_solver_iterations"][i]):
                first_flip = i
        final = abs(base["cost"][-1] - t["cost"][-1]) / abs(base["cost"][-1])
        rec = {"workload": base["workload"], "n_gpus": [base["n_gpus"], t["n_gpus"]], "trials": [len(base["cost"]), len(t["cost"])],
               "worst_step1": worst[1], "worst_step2": worst[2], "first_differing_trial": first_flip, "final": final}
        out.append(rec)
        print(json.dumps(rec))
    return out


if __name__ == "__main__":
    main()
