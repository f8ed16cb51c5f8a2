"""Do a wgrad and a dgrad kernel sharing the SMs (two streams, half the grid each) beat running them back to back?
Both read the same dZ [M,128]; run as: python profiles/concurrency_probe.py (spawns itself with PINNK_SM_COUNT)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    _R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
    import torch
    from pinns_rl_pde_b200 import _lib
    dev = torch.device('cuda:0')
    M, K, N, C = 4 * (1 << 17), 128, 128, 4
    X = torch.randn(M, K, device=dev); dZ = torch.randn(M, N, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_(); torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize()
            tot += a.elapsed_time(e)
        return tot / reps
    def serial():
        _lib.debug_linear_wgrad(dZ, X, C, 1); _lib.debug_linear_dgrad(dZ, W, 1)
    def wg(): _lib.debug_linear_wgrad(dZ, X, C, 1)
    def dg(): _lib.debug_linear_dgrad(dZ, W, 1)
    def conc():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1): _lib.debug_linear_wgrad(dZ, X, C, 1)
        with torch.cuda.stream(s2): _lib.debug_linear_dgrad(dZ, W, 1)
        cur.wait_stream(s1); cur.wait_stream(s2)
    print(f"SMs={sys.argv[1]}: wgrad {timed(wg):.3f} ms  dgrad {timed(dg):.3f} ms  serial {timed(serial):.3f} ms  two streams {timed(conc):.3f} ms", flush=True)
else:
    for c in (148, 74, 64, 84):
        env = dict(os.environ, PINNK_SM_COUNT=str(c))
        subprocess.run([sys.executable, __file__, str(c)], env=env)
