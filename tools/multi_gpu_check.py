"""Multi-GPU check of alabi_b200.parallel over NCCL (run with torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py

Rank 0 trains, L and alpha are broadcast, then every sharded operation is compared with
the same operation done by rank 0 alone on the whole input: predictions and the utility
argmin must be IDENTICAL (bits / index), restarts must pick the serial winner, and the
gathered sub-ensemble chain must equal the chains the ranks produced locally."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alabi_b200 as ab
from alabi_b200 import parallel as par
from alabi_b200.ensemble import EnsembleSampler, SurrogateLogProb

rank, world, local = par.init_distributed()
dev = torch.device("cuda", local)
rng = np.random.default_rng(123)                      # same stream on every rank
n, d = 1500, 4
X = rng.uniform(-2, 2, size=(n, d))
y = -0.5 * np.sum((X / 0.8) ** 2, axis=1) + 0.01 * rng.normal(size=n)
mk = lambda: ab.GP(kernel=ab.kernels.Matern52Kernel(metric=np.full(d, 1.3), ndim=d) * np.var(y), fit_mean=True,
                   mean=np.median(y), white_noise=-7.0, fit_white_noise=True)
gp = mk()
if rank == 0:
    gp.compute(X)
    gp._set_targets(y)
par.broadcast_gp(gp)
out = {"world": world}

# sharded predict == whole-set predict (bit-identical)
xq = torch.from_numpy(rng.uniform(-2, 2, size=(40001, d))).to(dev)
(mu_g, var_g), (lo, hi) = par.sharded_predict(gp, y, xq, return_var=True, gather=True)
mu_1, var_1 = gp.predict(y, xq, return_var=True)
out["predict_identical"] = bool(torch.equal(mu_g, mu_1) and torch.equal(var_g, var_1))

# sharded utility argmin == single-GPU argmin (same index), for every utility
b = np.array([(-2.0, 2.0)] * d)
cand = torch.from_numpy(rng.uniform(-2.2, 2.2, size=(30011, d))).to(dev)
for algo in ("bape", "agp", "jones"):
    v_s, i_s = par.sharded_utility_argmin(gp, y, cand, b, algorithm=algo, y_best=float(np.max(y)))
    i_1, v_1 = gp.utility_argmin(y, cand, b, algorithm=algo, y_best=float(np.max(y)))
    out[f"argmin_{algo}"] = bool(i_s == i_1 and v_s == v_1)

# sharded restarts pick the serial winner
from scipy.optimize import minimize
starts = par.broadcast_object(np.random.default_rng(7 + rank).uniform(-3, 3, size=(6, 2)))
fun = lambda x: float(np.sum((x ** 2 - 1.0) ** 2) + 0.1 * x[0])
best, allr = par.sharded_restarts(lambda x0: minimize(fun, x0, method="l-bfgs-b"), starts)
serial = [minimize(fun, s0, method="l-bfgs-b") for s0 in starts]
want = min(range(len(starts)), key=lambda i: (serial[i].fun, i))
out["restarts_winner"] = bool(best.restart == want and np.allclose(best.x, serial[want].x))

# sharded sub-ensembles: gathered chain == per-rank chains, disjoint Philox walker ids
lp = SurrogateLogProb(gp, y, b)
nw, steps = 64, 30
p0 = rng.uniform(-1, 1, size=(nw, d))
s, chain = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=5), p0, steps)
lo, hi = par.shard_range(nw, rank, world)
out["ensemble_gather"] = bool(chain.shape == (steps, nw, d) and np.array_equal(chain[:, lo:hi], s.get_chain()))
ref = EnsembleSampler(hi - lo, d, lp, seed=5)
ref.run_mcmc(p0[lo:hi], steps, walker_offset=lo)
out["ensemble_reproducible"] = bool(np.array_equal(ref.get_chain(), s.get_chain()))
# the fused gather (the sampler kernels store into every rank's buffer over NVLink, default with nccl)
# and the NCCL all_gather after the run deliver the same gathered chain on EVERY rank
if world > 1:
    s_u, chain_u = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=5), p0, steps, fused=False)
    out["ensemble_fused_equals_nccl_gather"] = bool(np.array_equal(chain, chain_u))
    # a larger, wide-unit ensemble with thinning, twice through the same cached peer buffers
    nw2 = 4736 * 2
    p2 = rng.uniform(-1, 1, size=(nw2, d))
    for rep_ in range(2):
        s_f, ch_f = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=11 + rep_), p2, 6, thin_by=2)
        s_n, ch_n = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=11 + rep_), p2, 6, thin_by=2, fused=False)
        out[f"ensemble_fused_wide_{rep_}"] = bool(ch_f.shape == (6, nw2, d) and np.array_equal(ch_f, ch_n)
                                                  and np.array_equal(s_f.get_log_prob(), s_n.get_log_prob()))
    # the same run cut into pieces (rows of piece b copied to the host while piece b + 1 samples)
    s_p, ch_p = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=12), p2, 6, thin_by=2, pieces=4)
    out["ensemble_fused_pieces"] = bool(np.array_equal(ch_p, ch_n) and np.array_equal(s_p.get_chain(), s_n.get_chain())
                                        and np.array_equal(s_p.get_log_prob(), s_n.get_log_prob()))
    # a third shape: the least recently used peer buffers are released collectively and new ones mapped
    s_3, ch_3 = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=5), p0, 7)
    s_4, ch_4 = par.sharded_ensemble(lambda k: EnsembleSampler(k, d, lp, seed=5), p0, 7, fused=False)
    out["ensemble_fused_after_eviction"] = bool(np.array_equal(ch_3, ch_4) and len(par.PeerChainBuffers._cache) == 2)

# the library's own NCCL communicator (ab_nccl_*): factor state handle to handle, all_gather of records
if world > 1:
    comm = par.NativeComm()
    gp2 = mk()
    if rank == 0:
        gp2.compute(X)
        gp2._set_targets(y)
    comm.broadcast_gp(gp2)
    mu_n, var_n = gp2.predict(y, xq[:5000], return_var=True)
    out["native_broadcast_identical"] = bool(torch.equal(mu_n, mu_1[:5000]) and torch.equal(var_n, var_1[:5000]))
    rec = torch.tensor([float(rank), 10.0 * rank + 1.0], dtype=torch.float64, device=dev)
    got = comm.allgather(rec)
    comm.sync()
    want = torch.tensor([[float(r), 10.0 * r + 1.0] for r in range(world)], dtype=torch.float64, device=dev)
    out["native_allgather"] = bool(torch.equal(got, want))
    comm.close()

flags = torch.tensor([float(all(v for k, v in out.items() if k != "world"))], device=dev)
if world > 1:
    torch.distributed.all_reduce(flags, op=torch.distributed.ReduceOp.MIN)
out["all_ranks_ok"] = bool(flags.item() == 1.0)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
sys.exit(0 if out["all_ranks_ok"] else 1)
