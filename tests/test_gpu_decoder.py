"""GPU tests of the ONet decoder: tcgen05 plumbing self-test, fp32 exact path (1e-4) and bf16 tensor-core path
against the reference-generated golden logits / oracle/model_ref, plus full 32^3 properties."""
import numpy as np
import pytest
import torch

from oracle import model_ref
from rfdnet_b200 import _lib, onet
from rfdnet_b200.synth import seeded_fill

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# bf16 operands through ten 256x256 layers: error budget relative to the logit scale (measured ~3e-3)
BF16_REL_TOL = 2e-2


def test_umma_selftest():
    lib = _lib.load()
    torch.manual_seed(0)
    A = torch.randn(128, 64, device=DEV)
    B = torch.randn(256, 64, device=DEV)
    D = torch.empty(128, 256, device=DEV)
    _lib.check(lib.rfd_umma_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "umma_selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    assert torch.allclose(D, ref, atol=1e-3, rtol=1e-4), float((D - ref).abs().max())


def test_umma_selftest_ts():
    """A operand from tensor memory (the layout the decoder's epilogue writes with tcgen05.st.16x128b)."""
    lib = _lib.load()
    torch.manual_seed(1)
    A = torch.randn(128, 64, device=DEV)
    B = torch.randn(256, 64, device=DEV)
    D = torch.empty(128, 256, device=DEV)
    _lib.check(lib.rfd_umma_selftest_ts(A.data_ptr(), B.data_ptr(), D.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "umma_selftest_ts")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ B.bfloat16().float().t()
    assert torch.allclose(D, ref, atol=1e-3, rtol=1e-4), float((D - ref).abs().max())


def _decoder(seed=31):
    dec = onet.DecoderCBatchNorm(dim=3, z_dim=32, c_dim=512).eval()
    seeded_fill(dec, seed)
    return dec


def test_grid_matches_reference(golden):
    g = onet.make_3d_grid(32, 1.1, DEV).cpu()
    assert np.array_equal(g[:32, 2].numpy(), golden["grid32_axis"])
    assert torch.equal(g, model_ref.make_3d_grid(32, 1.1))


def test_decoder_fp32_path_vs_golden(golden):
    dec = _decoder().to(DEV)
    grid = onet.make_3d_grid(32, 1.1, DEV)
    p = grid[torch.from_numpy(golden["dec_sel"]).long().to(DEV)].contiguous()
    c = torch.from_numpy(golden["dec_c"]).to(DEV)
    with torch.no_grad():
        out = dec.decode(p, torch.zeros(3, 32, device=DEV), c, precision="fp32")
        out_z = dec.decode(p.unsqueeze(0).expand(3, -1, -1).contiguous(), torch.from_numpy(golden["dec_z2"]).to(DEV), c,
                           precision="fp32")
    assert np.allclose(out.cpu().numpy(), golden["dec_logits"], atol=1e-4, rtol=1e-4)      # north_star: 1e-4
    assert np.allclose(out_z.cpu().numpy(), golden["dec_logits_z"], atol=1e-4, rtol=1e-4)


def test_decoder_bf16_tensor_core_path_vs_golden(golden):
    dec = _decoder().to(DEV)
    grid = onet.make_3d_grid(32, 1.1, DEV)
    p = grid[torch.from_numpy(golden["dec_sel"]).long().to(DEV)].contiguous()
    c = torch.from_numpy(golden["dec_c"]).to(DEV)
    with torch.no_grad():
        out = dec(p.unsqueeze(0).expand(3, -1, -1).contiguous(), torch.zeros(3, 32, device=DEV), c)  # module forward
    ref = golden["dec_logits"]
    err = np.abs(out.cpu().numpy() - ref).max()
    scale = np.abs(ref).max()
    print(f"bf16 decoder: max|err| {err:.4e}, logit scale {scale:.3f}, rel {err / scale:.3e}")
    assert err <= BF16_REL_TOL * max(1.0, scale)
    assert ((out.cpu().numpy() >= 0) == (ref >= 0)).mean() > 0.99  # occupancy decision (threshold logit(0.5) = 0)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 127), (3, 129), (5, 1000), (300, 256)])
def test_decoder_bf16_vs_fp32_ragged_sizes(B, T):
    dec = _decoder(seed=7).to(DEV)
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
    p = (torch.rand(B, T, 3, generator=g) - 0.5).to(DEV)
    c = torch.randn(B, 512, generator=g).to(DEV)
    z = (torch.randn(B, 32, generator=g) * 0.2).to(DEV)
    with torch.no_grad():
        a = dec.decode(p, z, c, precision="fp32")
        b = dec.decode(p, z, c, precision="bf16")
    scale = float(a.abs().max())
    assert float((a - b).abs().max()) <= BF16_REL_TOL * max(1.0, scale)
    if B <= 3:
        sd = {k: v.cpu() for k, v in dec.state_dict().items()}
        ref = model_ref.decoder(p.cpu(), z.cpu(), c.cpu(), sd)
        assert torch.allclose(a.cpu(), ref, atol=1e-4, rtol=1e-4)


def test_decoder_full_grid_properties():
    """config 4 shape at reduced object count: 8 objects x 32^3 shared lattice.  Size-independent checks:
    object independence (row b only depends on c[b]) and agreement with the fp32 path on a strided subset."""
    dec = _decoder(seed=9).to(DEV)
    grid = onet.make_3d_grid(32, 1.1, DEV)
    g = torch.Generator().manual_seed(1)
    c = torch.randn(8, 512, generator=g).to(DEV)
    z = torch.zeros(8, 32, device=DEV)
    with torch.no_grad():
        full = dec.decode(grid, z, c, precision="bf16")
        assert full.shape == (8, 32768) and torch.isfinite(full).all()
        perm = torch.tensor([3, 0, 7, 1, 2, 6, 5, 4], device=DEV)
        full_p = dec.decode(grid, z, c[perm].contiguous(), precision="bf16")
        assert torch.equal(full_p, full[perm])                       # per-object, bitwise deterministic
        sub = grid[::37].contiguous()
        ref = dec.decode(sub, z, c, precision="fp32")
    scale = float(ref.abs().max())
    assert float((full[:, ::37] - ref).abs().max()) <= BF16_REL_TOL * max(1.0, scale)


@pytest.mark.parametrize("B,T", [(3, 32768), (5, 1000), (1, 1), (2, 33)])
def test_occupancy_bits_exact(B, T):
    g = torch.Generator().manual_seed(B * 7 + T)
    logits = torch.randn(B, T, generator=g).to(DEV)
    logits[0, 0] = 0.0  # exactly on the threshold counts as occupied (>=)
    bits, counts = onet.occupancy_bits(logits, 0.0)
    occ = (logits >= 0.0).cpu().numpy()
    assert np.array_equal(counts.cpu().numpy(), occ.sum(1).astype(np.int32))
    pad = np.zeros((B, (T + 31) // 32 * 32), bool)
    pad[:, :T] = occ
    words = (pad.reshape(B, -1, 32) * (1 << np.arange(32, dtype=np.uint64))).sum(-1).astype(np.uint32)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), words)
