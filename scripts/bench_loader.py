"""Matrix-Market loading: the reference's read_suitsparse_matrix + CSC_2_CSR (compiled
unmodified in oracle/_ref) against the product loader sx_load_mtx_* (parallel parse).
CPU only.  usage: bench_loader.py [entries] [threads,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import sextans_b200 as sx

nz = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
threads = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, os.cpu_count()]
M = K = max(1000, nz // 20)
path = f"/tmp/sx_loader_bench_{nz}.mtx"
if not os.path.exists(path):
    rng = np.random.default_rng(1)
    r = rng.integers(1, M + 1, size=nz); c = rng.integers(1, K + 1, size=nz); v = rng.uniform(-1, 1, size=nz)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real general\n{M} {K} {nz}\n")
        np.savetxt(f, np.column_stack([r, c, v]), fmt="%d %d %.17g")
print(f"{path}: {os.path.getsize(path) / 1e6:.0f} MB, {nz} entries, M=K={M}")

def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), out

t_ref, ref = best(lambda: oracle.ref_load_csr(path), reps=2)
print(f"reference loader (fscanf + qsort + CSC_2_CSR, 1 thread): {t_ref:7.2f} s   {nz / t_ref / 1e6:6.2f} M entries/s")
for th in sorted(set(threads)):
    os.environ["SX_LOADER_THREADS"] = str(th)
    t, mine = best(lambda: sx.load_mtx(path, np.float32))
    same = mine[:3] == ref[:3] and all(np.array_equal(a, b) for a, b in zip(mine[3:], ref[3:6]))
    print(f"sx_load_mtx_f32, {th:2d} threads: {t:7.2f} s   {nz / t / 1e6:6.2f} M entries/s   x{t_ref / t:5.1f}   {'== reference CSR' if same else 'DIFFERS'}")
