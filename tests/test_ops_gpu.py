"""Each CUDA op of libaurora_b200.so (called through the C ABI) against the oracle / a plain PyTorch
fp32 reference of the same op, on seeded inputs.  Index arithmetic is bit-exact; floating point is
checked at bf16-operand tolerances stated per test."""

import hashlib
import json
import math
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import windows as W

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
WS0 = (2, 6, 12)
DEV = "cuda"


# ------------------------------------------------------------------------------------------------
# window index arithmetic: bit exact
# ------------------------------------------------------------------------------------------------
RES_CASES = [(4, 45, 90), (4, 4, 8), (4, 15, 30), (4, 38, 75), (4, 12, 24), (6, 13, 25), (4, 8, 16), (4, 2, 4),
             (4, 1, 2), (4, 90, 180), (4, 180, 360), (4, 150, 300)]


@pytest.mark.parametrize("res", RES_CASES)
@pytest.mark.parametrize("shifted", [False, True])
@pytest.mark.parametrize("warped", [True, False])
def test_device_index_map_bit_exact(res, shifted, warped):
    from aurora_b200 import cabi

    ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
    idx, grp = cabi.window_index_map(res, WS0, ss0, warped)
    ref_idx = W.window_gather_map(res, WS0, ss0)[0]
    np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)
    ref_grp = W.window_group_ids(res, WS0, ss0, warped)
    if ref_grp is not None:
        np.testing.assert_array_equal(grp.cpu().numpy(), ref_grp)


def test_device_index_map_matches_reference_checksums():
    """Production shapes against checksums computed from the reference itself (make_golden.py)."""
    from aurora_b200 import cabi

    hashes = json.loads((GOLD / "windows_hashes.json").read_text())
    for key, want in hashes.items():
        kind, dims, sh, _ = key.split("_")
        res = tuple(int(v) for v in dims.split("x"))
        ss0 = tuple(s // 2 for s in WS0) if sh == "s" else (0, 0, 0)
        idx, grp = cabi.window_index_map(res, WS0, ss0, True)
        got = idx.cpu().numpy().astype(np.int32) if kind == "idx" else grp.cpu().numpy()
        assert hashlib.sha256(got.tobytes()).hexdigest() == want, key


# ------------------------------------------------------------------------------------------------
# window attention
# ------------------------------------------------------------------------------------------------
def _attention_reference(qkv, pad_qkv, batch, res, ss0, heads):
    """fp32 restatement through the oracle's gather map (same math as oracle.swin_block's core)."""
    c, h, w = res
    l = c * h * w
    d = heads * 64
    idx_np, ws, ss, _ = W.window_gather_map(res, WS0, ss0)
    idx = torch.from_numpy(idx_np).to(qkv.device)
    nw, n = idx.shape
    valid = idx >= 0
    x = qkv.float().view(batch, l, 3 * d)
    xw = pad_qkv.float().view(1, 1, 1, 3 * d).expand(batch, nw, n, 3 * d).clone()
    xw[:, valid] = x[:, idx[valid]]
    q, k, v = xw.view(batch, nw, n, 3, heads, 64).permute(3, 0, 1, 4, 2, 5)
    logits = q @ k.transpose(-1, -2) / 8.0
    mask = W.shifted_window_mask(res, WS0, ss0, True)
    if mask is not None:
        logits = logits + torch.from_numpy(mask).to(qkv.device)[None, :, None]
    o = torch.softmax(logits, -1) @ v
    o = o.permute(0, 1, 3, 2, 4).reshape(batch, nw, n, d)
    out = torch.zeros(batch, l, d, device=qkv.device)
    out[:, idx[valid]] = o[:, valid]
    return out.view(batch * l, d)


@pytest.mark.parametrize("res,heads,batch", [((4, 12, 24), 2, 1), ((4, 15, 30), 4, 2), ((4, 8, 16), 2, 1),
                                             ((4, 4, 8), 8, 1), ((4, 2, 4), 4, 2), ((4, 1, 2), 2, 1),
                                             ((4, 45, 90), 4, 1)])
@pytest.mark.parametrize("shifted", [False, True])
def test_window_attention_matches_oracle(res, heads, batch, shifted):
    from aurora_b200 import cabi

    torch.manual_seed(hash((res, heads, shifted)) % 1000)
    ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
    l = res[0] * res[1] * res[2]
    d = heads * 64
    qkv = (torch.randn(batch * l, 3 * d, device=DEV) * 1.5).to(torch.bfloat16)
    pad = torch.randn(3 * d, device=DEV).to(torch.bfloat16)
    out = torch.full((batch * l, d), float("nan"), device=DEV, dtype=torch.bfloat16)
    cabi.window_attention(qkv, out, batch=batch, res=res, window=WS0, shift=ss0, num_heads=heads, pad_qkv=pad)
    torch.cuda.synchronize()
    ref = _attention_reference(qkv, pad, batch, res, ss0, heads)
    assert torch.isfinite(out.float()).all()
    # P is rounded to bf16 before P.V and the output is stored in bf16: ~2^-8 relative.
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    assert (out.float() - ref).abs().mean().item() < 3e-3


def test_window_attention_optional_bias():
    from aurora_b200 import cabi

    torch.manual_seed(1)
    res, heads = (4, 12, 24), 2
    l, d = 4 * 12 * 24, 128
    qkv = torch.randn(l, 3 * d, device=DEV).to(torch.bfloat16)
    bias = torch.randn(heads, 144, 144, device=DEV)
    out = torch.empty(l, d, device=DEV, dtype=torch.bfloat16)
    cabi.window_attention(qkv, out, batch=1, res=res, window=WS0, shift=(0, 0, 0), num_heads=heads, bias=bias)
    idx = torch.from_numpy(W.window_gather_map(res, WS0, (0, 0, 0))[0]).to(DEV)
    xw = qkv.float()[idx]
    q, k, v = xw.view(idx.shape[0], 144, 3, heads, 64).permute(2, 0, 3, 1, 4)
    o = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + bias[None], -1) @ v
    ref = torch.zeros(l, d, device=DEV)
    ref[idx] = o.permute(0, 2, 1, 3).reshape(idx.shape[0], 144, d)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------------
# LayerNorm + modulation + residual
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [128, 256, 512, 1024, 2048])
def test_ln_mod_residual(dim):
    from aurora_b200 import cabi

    torch.manual_seed(dim)
    rows = 777
    y = (torch.randn(rows, dim, device=DEV) * 3 + 1).to(torch.bfloat16)
    scale, shift = torch.randn(dim, device=DEV), torch.randn(dim, device=DEV)
    res = torch.randn(rows, dim, device=DEV)
    o32 = torch.empty(rows, dim, device=DEV)
    o16 = torch.empty(rows, dim, device=DEV, dtype=torch.bfloat16)
    cabi.ln_mod_residual(y, scale=scale, shift=shift, residual=res, out_f32=o32, out_bf16=o16)
    ref = res + F.layer_norm(y.float(), (dim,)) * scale + shift
    torch.testing.assert_close(o32, ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o16.float(), ref, rtol=1e-2, atol=1e-2)


def test_ln_mod_residual_broadcast_and_addrows_inplace():
    from aurora_b200 import cabi

    torch.manual_seed(3)
    nloc, lq, dim = 50, 3, 256
    rows = lq * nloc
    y = torch.randn(rows, dim, device=DEV).to(torch.bfloat16)
    gamma, beta = torch.randn(dim, device=DEV), torch.randn(dim, device=DEV)
    lat = torch.randn(lq, dim, device=DEV)
    add = torch.randn(nloc, dim, device=DEV)
    wide = torch.zeros(rows, 2 * dim, device=DEV, dtype=torch.bfloat16)
    o32 = torch.empty(rows, dim, device=DEV)
    cabi.ln_mod_residual(y, scale=gamma, shift=beta, residual=lat, res_div=nloc, res_mod=lq, add_rows=add,
                         out_f32=o32, out_bf16=wide[:, dim:], eps=1e-3)
    ref = lat.repeat_interleave(nloc, 0) + F.layer_norm(y.float(), (dim,), gamma, beta, 1e-3) + add.repeat(lq, 1)
    torch.testing.assert_close(o32, ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(wide[:, dim:].float(), ref, rtol=1e-2, atol=2e-2)
    assert wide[:, :dim].abs().max().item() == 0
    # in place on a strided bf16 view (the encoder's ln_k path)
    kv = torch.randn(rows, 2 * dim, device=DEV).to(torch.bfloat16)
    before = kv.clone()
    cabi.ln_mod_residual(kv[:, :dim], scale=gamma, shift=beta, out_bf16=kv[:, :dim])
    torch.testing.assert_close(kv[:, :dim].float(), F.layer_norm(before[:, :dim].float(), (dim,), gamma, beta),
                               rtol=1e-2, atol=2e-2)
    assert torch.equal(kv[:, dim:], before[:, dim:])


# ------------------------------------------------------------------------------------------------
# patch merging / splitting
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,h,w,d", [(4, 8, 16, 128), (4, 15, 30, 128), (2, 7, 9, 256), (4, 45, 90, 64)])
def test_patch_merge_ln(c, h, w, d):
    from aurora_b200 import cabi

    torch.manual_seed(h * w)
    x = torch.randn(1, c, h, w, d, device=DEV) * 2 + 0.5
    g, b = torch.randn(4 * d, device=DEV), torch.randn(4 * d, device=DEV)
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    out = torch.empty(c * h2 * w2, 4 * d, device=DEV, dtype=torch.bfloat16)
    cabi.patch_merge_ln(x, g, b, out, batch=1, c=c, h=h, w=w, d=d)
    xp = F.pad(x, (0, 0, 0, w % 2, 0, h % 2))
    m = xp.reshape(1, c, h2, 2, w2, 2, d).permute(0, 1, 2, 4, 3, 5, 6).reshape(c * h2 * w2, 4 * d)
    ref = F.layer_norm(m, (4 * d,), g, b)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("c,h,w,d,ch,cw", [(4, 4, 8, 256, 0, 0), (4, 8, 15, 256, 1, 0), (2, 4, 5, 128, 1, 1)])
def test_patch_split_ln(c, h, w, d, ch, cw):
    from aurora_b200 import cabi

    torch.manual_seed(h * w + d)
    y = torch.randn(c * h * w, 2 * d, device=DEV).to(torch.bfloat16)
    g, b = torch.randn(d // 2, device=DEV), torch.randn(d // 2, device=DEV)
    ho, wo = 2 * h - ch, 2 * w - cw
    out = torch.empty(c * ho * wo, d // 2, device=DEV, dtype=torch.bfloat16)
    cabi.patch_split_ln(y, g, b, out, batch=1, c=c, h=h, w=w, d=d, crop_h=ch, crop_w=cw)
    t = y.float().view(1, c, h, w, 2, 2, d // 2).permute(0, 1, 2, 4, 3, 5, 6).reshape(1, c, 2 * h, 2 * w, d // 2)
    t = t[:, :, :ho, :wo].reshape(-1, d // 2)
    ref = F.layer_norm(t, (d // 2,), g, b)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------------
# perceiver attention, small linear
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lq,lk,heads,dh", [(3, 13, 8, 32), (13, 3, 16, 64), (3, 4, 4, 32), (4, 3, 4, 64)])
def test_perceiver_attention(lq, lk, heads, dh):
    from aurora_b200 import cabi

    torch.manual_seed(lq * lk)
    nloc, d = 333, heads * dh
    q = torch.randn(lq, d, device=DEV)
    kv = torch.randn(lk * nloc, 2 * d, device=DEV).to(torch.bfloat16)
    out = torch.empty(lq * nloc, d, device=DEV, dtype=torch.bfloat16)
    cabi.perceiver_attention(q, kv, out, nloc=nloc, num_heads=heads, head_dim=dh)
    k, v = kv.float().view(lk, nloc, 2, heads, dh).permute(2, 1, 3, 0, 4)  # (nloc, heads, lk, dh)
    qh = q.view(lq, heads, dh).permute(1, 0, 2)[None]                      # (1, heads, lq, dh)
    o = torch.softmax(qh @ k.transpose(-1, -2) / math.sqrt(dh), -1) @ v    # (nloc, heads, lq, dh)
    ref = o.permute(2, 0, 1, 3).reshape(lq * nloc, d)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


def test_linear_small():
    from aurora_b200 import cabi

    torch.manual_seed(0)
    x, w, b = torch.randn(13, 512, device=DEV), torch.randn(1024, 512, device=DEV) / 22, torch.randn(1024, device=DEV)
    torch.testing.assert_close(cabi.linear_small(x, w, b), F.linear(x, w, b), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(cabi.linear_small(x, w, b, silu_in=True), F.linear(F.silu(x), w, b), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(cabi.linear_small(x, w, None, silu_out=True), F.silu(F.linear(x, w)), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# patchify / unpatchify
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p,h,w", [(4, 32, 64), (3, 45, 90), (10, 40, 80)])
def test_patchify_matches_im2col(p, h, w):
    from aurora_b200 import cabi

    torch.manual_seed(p)
    t, nv = 2, 5
    fields = [torch.randn(t, h, w, device=DEV) * (i + 1) + i for i in range(nv - 1)]
    static = torch.randn(h, w, device=DEV)
    descs = []
    for i, f in enumerate(fields):
        d = cabi.AbFieldIn()
        d.ptr, d.stride_t, d.loc, d.scale = f.data_ptr(), h * w, float(i), float(i + 1)
        d.transform = cabi.AB_IN_CLAMP_MIN0 if i == 1 else (cabi.AB_IN_CLAMP_LOG_COMBINE if i == 2 else 0)
        d.w0, d.w1, d.wb = 0.4, 0.6, 0.05
        descs.append(d)
    d = cabi.AbFieldIn()
    d.ptr, d.stride_t, d.loc, d.scale = static.data_ptr(), 0, 0.5, 2.0
    descs.append(d)
    c = cabi.AbFieldIn()
    c.ptr, c.const_value, c.scale = None, 0.25, 1.0
    descs.append(c)
    k = (nv + 1) * t * p * p
    kpad = (k + 63) // 64 * 64
    out = torch.zeros((h // p) * (w // p), kpad, device=DEV, dtype=torch.bfloat16)
    cabi.patchify(descs, t, h, w, p, out)
    planes = []
    for i, f in enumerate(fields):
        v = (f - float(i)) / float(i + 1)
        if i == 1:
            v = v.clamp(min=0)
        if i == 2:
            v = v.clamp(min=0)
            eps = 1e-4
            v = 0.4 * v.clamp(0, 2.5) + 0.6 * ((torch.log(v.clamp(min=eps)) - math.log(eps)) / -math.log(eps)) + 0.05
        planes.append(v)
    planes.append(((static - 0.5) / 2.0)[None].expand(t, -1, -1))
    planes.append(torch.full((t, h, w), 0.25, device=DEV))
    x = torch.stack(planes, 0)  # (V, T, H, W)
    ref = x.view(nv + 1, t, h // p, p, w // p, p).permute(2, 4, 0, 1, 3, 5).reshape((h // p) * (w // p), k)
    torch.testing.assert_close(out[:, :k].float(), ref, rtol=1e-2, atol=1e-2)
    if kpad > k:
        assert out[:, k:].abs().max().item() == 0


@pytest.mark.parametrize("p,h,w", [(4, 32, 64), (3, 45, 90)])
def test_unpatchify_matches_reference_layout(p, h, w):
    from aurora_b200 import cabi

    torch.manual_seed(p + 10)
    nv = 3
    l = (h // p) * (w // p)
    y = torch.randn(l, nv * p * p + 8, device=DEV)
    outs = [torch.full((h + 1, w), float("nan"), device=DEV) for _ in range(nv)]
    prev = torch.randn(h + 1, w, device=DEV) * 3 + 2
    descs = []
    for v in range(nv):
        d = cabi.AbFieldOut()
        d.ptr, d.loc, d.scale, d.col, d.mod_col = outs[v].data_ptr(), float(v), float(v + 2), v * p * p, -1
        descs.append(d)
    descs[1].mod_col, descs[1].prev = 2 * p * p, prev.data_ptr()
    descs[1].clamp_min0, descs[0].clamp_max1 = 1, 1
    cabi.unpatchify(descs[:2], y, h, w, p)
    img = y[:, : nv * p * p].view(h // p, w // p, nv, p, p).permute(2, 0, 3, 1, 4).reshape(nv, h, w)
    ref0 = img[0].clamp(max=1) * 2.0 + 0.0
    pn = (prev[:h] - 1.0) / 3.0
    ref1 = (img[1] + (1 + img[2]) * pn).clamp(min=0) * 3.0 + 1.0
    torch.testing.assert_close(outs[0][:h], ref0, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(outs[1][:h], ref1, rtol=1e-5, atol=1e-5)
    assert torch.isnan(outs[0][h]).all()


# ------------------------------------------------------------------------------------------------
# latitude slabs (multi-GPU sharding of one forecast), emulated on one GPU
# ------------------------------------------------------------------------------------------------
def _slab_views(full_qkv, res, d3, h0, rows, halo):
    c, h, w = res
    g = full_qkv.view(c, h, w, d3)
    idx_own = [(h0 + i) % h for i in range(rows)]
    idx_top = [(h0 - halo + i) % h for i in range(halo)]
    idx_bot = [(h0 + rows + i) % h for i in range(halo)]
    local = g[:, idx_own].contiguous().view(c * rows * w, d3)
    d = d3 // 3
    halo_t = torch.stack((g[:, idx_top], g[:, idx_bot]), 0)[..., d:].contiguous()  # [2, C, halo, W, 2D]: K | V only
    return local, halo_t, idx_own


@pytest.mark.parametrize("res,heads,splits", [((4, 24, 24), 2, (0, 12, 24)), ((4, 45, 90), 2, (0, 12, 24, 36, 45)),
                                              ((4, 36, 48), 4, (0, 8, 20, 36)),
                                              # 6-row bands of a 12-row grid: a foreign row is within reach BOTH ways
                                              # round the cyclic axis; kernel and halo_needs must pick the same side
                                              ((4, 12, 24), 2, (0, 6, 12))])
@pytest.mark.parametrize("shifted", [False, True])
def test_window_attention_latitude_slabs_equal_whole_grid(res, heads, splits, shifted):
    """Every rank's slab result must equal the corresponding rows of the whole-grid result bit for bit."""
    from aurora_b200 import cabi

    torch.manual_seed(5)
    ss0 = tuple(s // 2 for s in WS0) if shifted else (0, 0, 0)
    c, h, w = res
    l, d = c * h * w, heads * 64
    qkv = torch.randn(l, 3 * d, device=DEV).to(torch.bfloat16)
    pad = torch.randn(3 * d, device=DEV).to(torch.bfloat16)
    full = torch.empty(l, d, device=DEV, dtype=torch.bfloat16)
    cabi.window_attention(qkv, full, batch=1, res=res, window=WS0, shift=ss0, num_heads=heads, pad_qkv=pad)
    full_g = full.view(c, h, w, d)
    halo = 5
    for r in range(len(splits) - 1):
        h0, rows = splits[r], splits[r + 1] - splits[r]
        local, halo_t, idx_own = _slab_views(qkv, res, 3 * d, h0, rows, halo)
        # only the rows `halo_needs` names may be read: poison the others (a neighbour does not send them)
        from aurora_b200 import sharding

        above, below = sharding.halo_needs(h, WS0[1], ss0[1], h0, rows)
        halo_t[0, :, : halo - above] = float("nan")
        halo_t[1, :, below:] = float("nan")
        out = torch.full((c * rows * w, d), float("nan"), device=DEV, dtype=torch.bfloat16)
        cabi.window_attention(local, out, batch=1, res=res, window=WS0, shift=ss0, num_heads=heads, pad_qkv=pad,
                              slab=(h0, rows), halo_kv=halo_t)
        torch.cuda.synchronize()
        want = full_g[:, idx_own].reshape(c * rows * w, d)
        assert torch.equal(out, want), (r, (out.float() - want.float()).abs().max().item())
