"""GPU parity: tcgen05 GEMM and the ViT-B/32 forward vs the fp32 torch oracle (A9)."""
import numpy as np
import pytest
import torch

from oracle import hmsg_oracle as O
from holoagent_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("two_sm", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 768), (1000, 768, 3072), (65, 512, 768), (5000, 2304, 768)])
def test_gemm_tcgen05(engine, M, N, K, two_sm):
    engine.set_option("gemm_2sm", two_sm)
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).half()
    W = (torch.randn(N, K, generator=g) * 0.05).half()
    ref = A.float() @ W.float().T
    Ad, Wd = A.cuda(), W.cuda()
    Cd = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    engine.gemm_debug(Ad, Wd, Cd, M, N, K)
    out = Cd.cpu()
    engine.set_option("gemm_2sm", 1)
    err = (out - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item() + 1e-5, err


@pytest.fixture(scope="module")
def vit(engine):
    sd = synth.make_vit_weights()
    engine.encoder_load(sd)
    return sd


@pytest.mark.parametrize("B", [1, 5, 67])
def test_vit_forward(engine, vit, B):
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 3, 224, 224, generator=g) * 1.2
    ref = O.get_img_feats_batch_tensor(vit, x)
    out = engine.encode_images(x.numpy())
    # contract: embeddings within 1e-3 relative (of the unit norm) per component; fp16 operands /
    # fp32 accumulate typically give ~3e-4
    err = np.abs(out - ref).max()
    cos = np.sum(out * ref, axis=-1)
    print("max abs err", err, "min cos", cos.min())
    assert err <= 1e-3
    assert np.all(cos > 1 - 1e-5)
    assert np.allclose(np.linalg.norm(out, axis=-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("opt,val", [("gemm_2sm", 0), ("attn_variant", 1), ("attn_variant", 2), ("attn_variant", 3), ("attn_variant", 4)])
def test_vit_variants_agree(engine, vit, opt, val):
    x = torch.randn(70, 3, 224, 224, generator=torch.Generator().manual_seed(5))
    base = engine.encode_images(x.numpy())
    engine.set_option(opt, val)
    try:
        alt = engine.encode_images(x.numpy())
    finally:
        engine.set_option(opt, 1 if opt == "gemm_2sm" else 0)
    assert np.abs(alt - base).max() < 2e-4


def test_vit_device_path_matches_host_path(engine, vit):
    x = torch.randn(9, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    a = engine.encode_images(x.numpy())
    b = engine.encode_images(x.cuda()).cpu().numpy()
    assert np.array_equal(a, b)
