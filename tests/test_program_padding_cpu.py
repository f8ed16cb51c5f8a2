"""CPU: the width-padding bookkeeping of program.compile_network (zero-padded shadow parameters, padded / real flat-gradient
layouts) -- no kernels involved."""
import torch


def test_shadow_parameters_and_gradient_unpadding():
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import program
    torch.manual_seed(0)
    m = pk.make_model("siren", 2, 124, 4, torch.device("cpu"), omega_0=30.0)
    pr = program.compile_network(m)
    assert pr.padded and pr.grad_floats == sum(p.numel() for p in m.parameters())
    assert [o.out_dim for o in pr.ops if o.kind == 1] == [128, 128, 128, 128, 1]
    assert pr.pad_floats == sum(sh.numel() for _, sh, _, _ in pr.pad_entries)
    # the library reads the shadows: leading block = the real parameter, the rest exactly zero
    pr.refresh_shadows()
    for real, sh, _, _ in pr.pad_entries:
        blk = sh[:real.shape[0], :real.shape[1]] if real.dim() == 2 else sh[:real.shape[0]]
        assert torch.equal(blk, real.detach())
        rest = sh.clone()
        if real.dim() == 2:
            rest[:real.shape[0], :real.shape[1]] = 0
        else:
            rest[:real.shape[0]] = 0
        assert float(rest.abs().max()) == 0.0
    # a parameter update is picked up by the next refresh
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    pr.refresh_shadows()
    real, sh, _, _ = pr.pad_entries[0]
    assert torch.equal(sh[:real.shape[0], :real.shape[1]], real.detach())
    # unpadding: real entries of a padded flat gradient land at model.parameters() offsets, padded entries are dropped
    flat_pad = torch.randn(pr.pad_floats)
    flat = torch.zeros(pr.grad_floats)
    pr.unpad_add(flat_pad, flat)
    for (real, sh, off, poff), g in zip(pr.pad_entries, pr.split_flat(flat)):
        blk = flat_pad[poff:poff + sh.numel()].view_as(sh)
        want = blk[:real.shape[0], :real.shape[1]] if real.dim() == 2 else blk[:real.shape[0]]
        assert torch.equal(g, want)
    pr.unpad_add(flat_pad, flat)                       # accumulates
    assert torch.allclose(pr.split_flat(flat)[2], 2 * flat_pad[pr.pad_entries[2][3]:pr.pad_entries[2][3] + 128 * 128].view(128, 128)[:124, :124])


def test_widths_that_stay_unpadded(monkeypatch):
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import program
    dev = torch.device("cpu")
    assert not program.compile_network(pk.make_model("feedforward", 2, 128, 3, dev)).padded          # already on the tiles
    assert not program.compile_network(pk.make_model("feedforward", 2, 32, 3, dev)).padded           # 4x the work: stays exact fp32
    assert not program.compile_network(pk.make_model("resnet", 2, 124, 2, dev, num_blocks=2)).padded  # LayerNorm statistics would change
    assert program.compile_network(pk.make_model("feedforward", 2, 248, 3, dev)).padded
    monkeypatch.setenv("PINNK_DISABLE_PAD", "1")
    assert not program.compile_network(pk.make_model("feedforward", 2, 248, 3, dev)).padded


def test_shadows_follow_the_model_to_another_device_or_dtype():
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import program
    m = pk.make_model("siren", 2, 124, 3, torch.device("cpu"), omega_0=30.0)
    pr = program.compile_network(m)
    m.double()                                  # same Parameter objects, new storage (what model.to(device) does as well)
    pr.refresh_shadows()
    for real, sh, _, _ in pr.pad_entries:
        assert sh.dtype == real.dtype and sh.device == real.device
        blk = sh[:real.shape[0], :real.shape[1]] if real.dim() == 2 else sh[:real.shape[0]]
        assert torch.equal(blk, real.detach())
