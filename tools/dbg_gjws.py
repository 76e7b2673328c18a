import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import oracle as orc
    from cuda_matrix_inversion_b200 import api
    mode = sys.argv[1]
    n = 128
    rng = np.random.default_rng(0)
    if mode == "diag":
        a = (rng.random((2, n, n)) * 0.01 + np.eye(n) * np.arange(n + 1, 1, -1)).astype(np.float32)   # pivots on the diagonal, in order
    else:
        a = (rng.random((2, n, n)) + np.eye(n)).astype(np.float32)
    flat = orc.to_colmajor(a)
    got, info = api.general_inverse_host(flat, n)
    g = orc.from_colmajor(got, n).astype(np.float64)
    ex = np.linalg.inv(a.astype(np.float64))
    err = np.abs(g - ex) / np.abs(ex).max()
    print(mode, os.environ.get("INVGPU_GJR2_WS"), "info", info, "max err", err.max())
    e = err[0]
    print(" per col-block of 32 (max):", [float(f"{e[:, 32*i:32*i+32].max():.2e}") for i in range(4)])
    print(" per row-block of 32 (max):", [float(f"{e[32*i:32*i+32].max():.2e}") for i in range(4)])
    np.set_printoptions(precision=4, linewidth=200)
    print(" diag got/ex:", (np.diag(g[0]) / np.diag(ex[0]))[[0,1,2,3,30,31,32,33,62,63,64,65,126,127]])
    print(" got[0][:4,:4] / ex:", (g[0][:4,:4] / ex[0][:4,:4]))
    print(" got[0][124:,124:] / ex:", (g[0][124:,124:] / ex[0][124:,124:]))
    print(" first bad column:", int(np.argmax(e.max(0) > 1e-3)), "first bad row:", int(np.argmax(e.max(1) > 1e-3)))
else:
    for mode in ("diag", "rand"):
        for ws in ("1", "0"):
            subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, INVGPU_GJR2_WS=ws))
