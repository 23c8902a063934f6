"""One GPU's share of BASELINE config 5 (N = 1e8, D = 256 over 8 GPUs -> 1.25e7 rows per GPU), rows generated on the
device (bnuts_model_logistic_synthetic), 4096 chains: time of the full-size gradient launch of k_logistic_tc256 and of a
lockstep leapfrog step, with the two-term residual and with the single-term residual the 8-GPU group takes by itself
(N_total >= 3.3e5 D; forced here with BNUTS_TC_RREF=1 because one shard alone is below the threshold).
Run under gpurun; prints one JSON line per mode."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import inplacedhmc_jl_b200 as bn

N = int(os.environ.get("NROWS", 12_500_000)); D = int(os.environ.get("DIM", 256)); C = int(os.environ.get("CHAINS", 4096))
ROW0 = int(os.environ.get("ROW0", 3 * 12_500_000))          # the shard of rank 3 of 8
SEED = 5
t0 = time.perf_counter()
e = bn.Engine(C, D, dtype=bn.F32, gradient_path=bn.GRAD_TENSOR)
e.model_logistic_synthetic(SEED, ROW0, N, 1.0)
torch.cuda.synchronize(); t_gen = time.perf_counter() - t0
_, _, beta = bn.synth_logistic_rows(SEED, 0, 0, D)
rng = np.random.default_rng(1)
e.set_positions(np.repeat(beta[None, :], C, axis=0))
t0 = time.perf_counter(); e.find_local_optimum(1e-4, int(os.environ.get("OPT_ITERS", 30))); t_opt = time.perf_counter() - t0
b = e.get_state()[0].mean(axis=0)
sd = 2.0 / np.sqrt(N)
q = b[None, :] + rng.normal(size=(C, D)) * sd
p = rng.normal(size=(C, D)) * np.sqrt(N) * 0.3
res = {}
for mode in ("0", "1"):
    os.environ["BNUTS_TC_RREF"] = mode
    e.logistic_set_reference(b)
    e.set_positions(q)
    res[mode] = e.get_state()[1]
    e.leapfrog(p, 1e-4, 2)
    walls = []
    e.profile(True)
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); e.leapfrog(p, 1e-4, 4); torch.cuda.synchronize()
        walls.append((time.perf_counter() - t) / 4)
    ms, n = e.profile(False)
    fl = 4.0 * N * D * C
    print(json.dumps({"config": "c5 shard: rows [%d, %d) of N=1e8, D=%d, %d chains, 1 GPU" % (ROW0, ROW0 + N, D, C),
                      "residual_terms": 2 if mode == "0" else 1, "generate_s": t_gen, "find_local_optimum_s": t_opt,
                      "grad_kernel_ms": ms / n, "alg_TFLOPs": fl / (ms / n * 1e-3) / 1e12,
                      "frac_of_sustained_bf16_peak_1390": fl / (ms / n * 1e-3) / 1e12 / 1390.3,
                      "lockstep_step_ms_min_of_3_calls": min(walls) * 1e3, "lockstep_step_ms_all": [w * 1e3 for w in walls],
                      "chain_leapfrogs_per_s": C / min(walls),
                      "grad_norm_mean": float(np.linalg.norm(res[mode], axis=1).mean())}), flush=True)
print(json.dumps({"single_vs_two_term_rel_diff_max": float(np.max(np.linalg.norm(res["1"] - res["0"], axis=1) / np.linalg.norm(res["0"], axis=1))),
                  "note": "same positions (about one posterior sd from the reference); the single-term error model is 1.7e-3 sqrt(D/N_total)"}))
