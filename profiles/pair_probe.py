"""Where does the paired reverse kernel spend its time?  One 4 Mi-row Linear(128,128)+tanh reverse step:
  * two launches on 148 SMs (the baseline), each role alone on 148 and on 74 SMs (per-SM headroom when HBM is uncontended),
  * the paired kernel with the lock-step lead swept (PINNK_PAIR_LEAD) up to "no throttle".
Spawns itself per setting (the library reads PINNK_SM_COUNT per call, the lead per launch)."""
import os, sys, subprocess
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
if len(sys.argv) > 1:
    import torch
    from pinns_rl_pde_b200 import _lib
    dev = torch.device('cuda:0')
    P, JC = 1 << 20, 4
    M = P * JC
    g = torch.Generator(device='cuda').manual_seed(0)
    Y = torch.empty(P, JC, 128, device=dev)
    Y[:, 0] = torch.tanh(torch.randn(P, 128, generator=g, device=dev)); Y[:, 1:] = torch.randn(P, JC - 1, 128, generator=g, device=dev) * 0.3
    Y = Y.reshape(M, 128).contiguous()
    dZ = torch.randn(M, 128, generator=g, device=dev); W = torch.randn(128, 128, generator=g, device=dev) / 11.3
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    def timed(fn, reps=6):
        for _ in range(2): fn()
        torch.cuda.synchronize(); tot = 0.0
        for _ in range(reps):
            flush.zero_(); torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); torch.cuda.synchronize(); tot += a.elapsed_time(e)
        return tot / reps
    mode = sys.argv[1]
    if mode == "split":
        t = timed(lambda: _lib.debug_bwd_layer(dZ, W, Y, 2, 1, False))
    elif mode == "pair":
        t = timed(lambda: _lib.debug_bwd_layer(dZ, W, Y, 2, 1, True))
    elif mode == "wgrad":
        t = timed(lambda: _lib.debug_linear_wgrad(dZ, Y, JC, 1))
    elif mode == "dgrad":       # plain dgrad (no adjoint): the rows kernel without the Y reads
        t = timed(lambda: _lib.debug_linear_dgrad(dZ, W, 1))
    print(f"{mode:6s} SMs={os.environ.get('PINNK_SM_COUNT','148'):>4s} lead={os.environ.get('PINNK_PAIR_LEAD','4'):>6s}: {t:.3f} ms", flush=True)
else:
    def run(mode, **env):
        subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, **{k: str(v) for k, v in env.items()}))
    run("split"); run("wgrad"); run("dgrad")
    run("split", PINNK_SM_COUNT=74); run("wgrad", PINNK_SM_COUNT=74); run("dgrad", PINNK_SM_COUNT=74)
    for lead in (2, 4, 8, 16, 64, 100000):
        run("pair", PINNK_PAIR_LEAD=lead)
    run("pair", PINNK_SM_COUNT=74)
