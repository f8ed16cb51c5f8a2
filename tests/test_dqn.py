"""Q-network scoring of the RL sampler (rl/rl_agent.py:15-88,214-229).
CPU: the oracle port against the fixture generated from the unmodified reference, host-side lowering and errors.
GPU: pinnk_dqn_forward against the fixture, against the oracle at other shapes, and with live dropout against the same
torch modules on the same CUDA generator state."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "x_dqn.npz")


def _fixture():
    z = np.load(GOLD)
    state = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    return state, torch.from_numpy(z["states"]), torch.from_numpy(z["q32"]), torch.from_numpy(z["q64"])


def test_oracle_port_matches_reference_fixture():
    state, pts, q32, q64 = _fixture()
    assert torch.equal(ref_port.dqn_forward_port(state, pts), q32)
    s64 = {k: v.double() for k, v in state.items()}
    assert torch.allclose(ref_port.dqn_forward_port(s64, pts.double()), q64, rtol=0, atol=1e-14)


def test_mirror_loads_reference_state_dict_and_refuses_cpu():
    from pinns_rl_pde_b200 import rl, _lib
    state, pts, _, _ = _fixture()
    net = rl.DQNNetwork(2, 1, 128)
    net.load_state_dict(state)                       # same keys and shapes as the reference's DQNNetwork
    groups, out = rl._lower(net)
    assert len(groups) == 2 and out.out_features == 1 and groups[0][2] == pytest.approx(0.1)
    with pytest.raises(_lib.PinnkError):
        rl.dqn_forward(net, pts)                     # CPU tensors: no fallback
    with pytest.raises(rl.UnsupportedQNetwork):
        rl._lower(torch.nn.Sequential(torch.nn.Linear(2, 1)))
    rl._lower(rl.DQNNetwork(2, 1, 512))               # the shipped width (config.yaml:363): the tcgen05 route takes it
    with pytest.raises(rl.UnsupportedQNetwork):       # neither <= 128 nor a multiple of 128: declined, loudly
        rl._lower(rl.DQNNetwork(2, 1, 200))


def _torch_twin(net):
    """The same network as plain torch modules (what the reference executes), sharing the parameters."""
    return torch.nn.Sequential(*[torch.nn.Sequential(*list(g)) if isinstance(g, torch.nn.Sequential) else g
                                 for g in net.layers])


@pytest.mark.gpu
def test_gpu_matches_reference_fixture():
    from pinns_rl_pde_b200 import rl, _lib
    dev = torch.device("cuda:0")
    state, pts, q32, q64 = _fixture()
    net = rl.DQNNetwork(2, 1, 128).to(dev).eval()
    net.load_state_dict(state)
    before = _lib.launch_count()
    q = net(pts.to(dev))
    assert _lib.launch_count() == before + 1
    err = float((q.double().cpu() - q64).norm() / q64.norm())
    floor = float((q32.double() - q64).norm() / q64.norm())
    assert err <= max(1e-5, 2 * floor), (err, floor)


@pytest.mark.gpu
@pytest.mark.parametrize("state_dim,hidden,layers,actions,n", [(2, 128, 3, 1, 10000), (3, 64, 4, 1, 1003), (2, 256, 2, 4, 77),
                                                              (4, 100, 3, 2, 1), (2, 512, 3, 1, 513), (5, 50, 3, 3, 40)])
def test_gpu_shapes_against_oracle(state_dim, hidden, layers, actions, n, monkeypatch):
    from pinns_rl_pde_b200 import rl
    monkeypatch.setattr(rl, "MAX_HIDDEN", 1024)          # the kernel itself covers what fits in shared memory
    dev = torch.device("cuda:0")
    torch.manual_seed(state_dim * 1000 + hidden)
    net = rl.DQNNetwork(state_dim, actions, hidden, num_layers=layers).to(dev).eval()
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("bias") or ".1.weight" in name:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.rand(n, state_dim, device=dev) * 2 - 1
    q = rl.dqn_forward(net, x)
    assert q.shape == (n, actions)
    s64 = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
    ref = ref_port.dqn_forward_port(s64, x.double().cpu())
    err = float((q.double().cpu() - ref).norm() / ref.norm())
    assert err <= 1e-5, err
    assert rl.dqn_forward(net, x[:0]).shape == (0, actions)


@pytest.mark.gpu
@pytest.mark.parametrize("state_dim,hidden,layers,actions,n,train", [
    (2, 512, 3, 1, 10000, False),      # the shipped agent (config.yaml:363, train.py:348-351) on the 100 x 100 grid
    (2, 512, 3, 1, 10000, True),       # with live dropout, as the reference scores (policy_net stays in train mode)
    (2, 256, 3, 1, 4097, False), (3, 384, 4, 2, 1000, True), (2, 1024, 2, 1, 333, False), (8, 640, 3, 1, 64, False),
    (2, 512, 3, 1, 1, False), (2, 512, 3, 1, 1000000, False)])
def test_gpu_wide_q_networks_on_the_tensor_core_route(state_dim, hidden, layers, actions, n, train):
    """pinnk_dqn_forward_wide: hidden Linear layers on the tcgen05 3xTF32 rows kernel, LayerNorm / ReLU / mask / first and
    output layer in one row kernel -- against the same torch modules in fp64 (eval) or on the same generator state (train)."""
    import parity_log
    from pinns_rl_pde_b200 import rl, _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(hidden + n)
    net = rl.DQNNetwork(state_dim, actions, hidden, num_layers=layers).to(dev)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("bias") or ".1.weight" in name:
                p.add_(0.1 * torch.randn_like(p))
    net.train(train)
    twin = _torch_twin(net).train(train)
    x = torch.rand(n, state_dim, device=dev) * 2 - 1
    torch.manual_seed(5)
    before = _lib.launch_count()
    q = rl.dqn_forward(net, x)
    launches = _lib.launch_count() - before
    after = torch.rand(4, device=dev)
    assert q.shape == (n, actions) and launches == (layers - 1) + (layers - 2) * (hidden // 128)
    torch.manual_seed(5)
    with torch.no_grad():
        ref32 = twin(x)
        after_ref = torch.rand(4, device=dev)
        if train:
            ref = ref32.double()                                 # same masks only through the same generator calls
        else:
            import copy
            ref = copy.deepcopy(twin).double()(x.double())
    assert torch.equal(after, after_ref)                         # the generator advanced exactly as under the torch modules
    err = float((q.double() - ref).norm() / ref.norm())
    floor = float((ref32.double() - ref).norm() / ref.norm())
    parity_log.log(f"[dqn wide] hidden {hidden} x {layers - 1} groups, {n} states, {'train' if train else 'eval'}: rel err vs "
                   f"{'torch fp32 (same masks)' if train else 'fp64'} {err:.2e} (torch fp32 vs fp64 {floor:.2e}), {launches} launches")
    assert err <= max(1e-5, 2 * floor), (err, floor)


@pytest.mark.gpu
def test_gpu_live_dropout_uses_the_reference_rng_stream():
    """policy_net stays in train mode in the reference, so its dropout is live during scoring: same CUDA generator state ->
    same masks as the torch modules, and the generator ends in the same state."""
    from pinns_rl_pde_b200 import rl
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    net = rl.DQNNetwork(2, 1, 128).to(dev).train()
    twin = _torch_twin(net).train()
    x = torch.rand(10000, 2, device=dev)
    torch.manual_seed(11)
    with torch.no_grad():
        ref = twin(x)
    after_ref = torch.rand(4, device=dev)
    torch.manual_seed(11)
    q = rl.dqn_forward(net, x)
    after = torch.rand(4, device=dev)
    assert torch.equal(after, after_ref)
    err = float((q - ref).norm() / ref.norm())
    assert err <= 1e-5, err
    assert float((q - rl.dqn_forward(net.eval(), x)).abs().max()) > 1e-3      # the masks did something


@pytest.mark.gpu
def test_adaptive_sampling_goes_through_libpinnk():
    """generate_collocation_points('adaptive') (pde_base.py:961-1072) with an agent holding a DQNNetwork: the grid scores come
    from pinnk_dqn_forward and the draw equals the one made from the torch modules' scores under the same seeds."""
    import types
    from helpers import product_pde
    from pinns_rl_pde_b200 import rl, _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = rl.DQNNetwork(2, 1, 128).to(dev).eval()
    agent = types.SimpleNamespace(policy_net=net, epsilon=-1.0, device=dev)
    twin = _torch_twin(net).eval()
    agent.select_action = lambda s: twin(s).view(1, -1)
    pde = product_pde("burgers", dev)
    pde.rl_agent = agent
    torch.manual_seed(7)
    before = _lib.launch_count()
    x, t = pde.generate_collocation_points(400, strategy="adaptive")
    assert _lib.launch_count() == before + 1
    assert x.shape == (400, 1) and t.shape == (400, 1)
    assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0 and float(t.min()) >= 0.0 and float(t.max()) <= 1.0
    pts = pde.collocation_history[-1]
    gs = 20
    axes = [torch.linspace(-1, 1, gs, device=dev), torch.linspace(0, 1, gs, device=dev)]
    grid = torch.stack([g.flatten() for g in torch.meshgrid(*axes, indexing="ij")], dim=1)
    p_lib = rl.grid_scores(agent, grid)
    with torch.no_grad():
        p_ref = torch.abs(twin(grid).view(1, -1))
        p_ref = p_ref / p_ref.sum()
    assert torch.allclose(p_lib, p_ref, rtol=1e-4, atol=1e-9)
    assert pts.shape == (400, 2)


def test_adaptive_sampler_keeps_foreign_agents_on_their_own_forward():
    """An agent without a DQNNetwork-shaped ``policy_net`` (or a wider one than the kernel is used for) is not replaced:
    grid_scores falls through to ``agent.select_action`` (pde_base.py:1003-1008 unchanged).  Runs on the CPU."""
    import types
    from helpers import product_pde
    from pinns_rl_pde_b200 import rl
    calls = []

    def select_action(points):
        calls.append(points.shape)
        return torch.linspace(1.0, 2.0, points.shape[0]).view(1, -1)

    agent = types.SimpleNamespace(select_action=select_action, epsilon=0.0, device=torch.device("cpu"))
    pts = torch.rand(50, 2)
    p = rl.grid_scores(agent, pts)
    assert calls == [pts.shape] and abs(float(p.sum()) - 1.0) < 1e-6 and p.shape == (1, 50)
    agent.policy_net = rl.DQNNetwork(2, 1, 200)                       # wider than rl.MAX_HIDDEN and not a multiple of 128: declined, agent's own forward
    rl.grid_scores(agent, pts)
    assert len(calls) == 2
    pde = product_pde("burgers", torch.device("cpu"))
    pde.rl_agent = agent
    torch.manual_seed(0)
    x, t = pde.generate_collocation_points(300, strategy="adaptive")
    assert x.shape == (300, 1) and t.shape == (300, 1) and len(calls) == 3
    assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0 and float(t.min()) >= 0.0 and float(t.max()) <= 1.0
