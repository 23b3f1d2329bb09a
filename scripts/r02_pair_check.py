"""Quick stand-alone check of the cta_group::2 pair scan (run on the GPU box before the full suite)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import dense as dense_oracle
from sentio_b200.engine import B200Engine

eng = B200Engine(0)
rng = np.random.default_rng(0)
for (n, d, B, k) in [(20000, 256, 128, 10), (50000, 1024, 200, 100), (9000, 64, 65, 5)]:
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x16 = x.astype(np.float16)
    q = rng.standard_normal((B, d)).astype(np.float32)
    eng.load_dense(x16)
    t = time.time()
    ids, sc, cnt = eng.dense_topk(q, k)
    dt = time.time() - t
    bad = 0
    for b in range(B):
        wi, ws = dense_oracle.dense_topk(x16, q[b], k)
        if list(ids[b]) != list(wi) or not np.allclose(sc[b], ws, rtol=1e-9, atol=1e-12):
            bad += 1
    print(f"n={n} d={d} B={B} k={k}: {bad} mismatching queries, {dt*1e3:.1f} ms", flush=True)
    assert bad == 0
print("pair check ok")
