"""HBM-bound kernels (pool+concat, L2 normalise, pack, negative gather) vs the oracle / golden vectors."""
import numpy as np
import pytest
import torch

from oracle import spec

pytestmark = pytest.mark.gpu


def cu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t.to(dtype) if dtype is not None else t


def test_pool_concat_golden(golden):
    from vast_b200 import ops
    g = golden("features")
    out = ops.pool_concat(cu(g["vision"]), cu(g["audio"]), cu(g["subtitle"]), vision_mode=0, audio_mode=1)
    ref = np.concatenate([g["pool_v"], g["pool_a"], g["pool_s"]], axis=1)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    out = ops.pool_concat(cu(g["vision"]), cu(g["audio"]), None, vision_mode=1, audio_mode=0)
    ref = np.concatenate([g["pool_v_swin"], g["pool_a_ast"]], axis=1)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_pool_concat_shapes(dtype):
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(3)
    bs = 9
    vis = torch.randn(bs, 4, 257, 1408, generator=gen).to(dtype)
    aud = torch.randn(bs, 2, 256, 768, generator=gen).to(dtype)
    sub = torch.randn(bs, 70, 768, generator=gen).to(dtype)
    out = ops.pool_concat(vis.cuda(), aud.cuda(), sub.cuda(), out_dtype=torch.float32).cpu().numpy()
    ref = np.concatenate([spec.pool_vision_for_contra(vis.float().numpy(), "evaclip"),
                          spec.pool_audio_for_contra(aud.float().numpy(), "beats"),
                          spec.pool_text_for_contra(sub.float().numpy())], axis=1)
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)
    # odd channel count -> scalar path
    v2 = torch.randn(3, 2, 5, 37, generator=gen).to(dtype)
    out = ops.pool_concat(v2.cuda(), None, None, vision_mode=1, out_dtype=torch.float32).cpu().numpy()
    np.testing.assert_allclose(out, spec.pool_vision_for_contra(v2.float().numpy(), "swin"), rtol=2e-5, atol=2e-6)


def test_pool_concat_bwd_matches_autograd():
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(4)
    vis = torch.randn(3, 2, 5, 16, generator=gen, requires_grad=True)
    aud = torch.randn(3, 2, 4, 8, generator=gen, requires_grad=True)
    sub = torch.randn(3, 6, 8, generator=gen, requires_grad=True)
    out = torch.cat([vis[:, :, 0].mean(1), aud.mean(2).mean(1), sub[:, 0]], dim=1)
    go = torch.randn(out.shape, generator=gen)
    out.backward(go)
    gv, ga, gs = ops.pool_concat_bwd(go.cuda(), tuple(vis.shape), tuple(aud.shape), tuple(sub.shape), 0, 1)
    np.testing.assert_allclose(gv.cpu().numpy(), vis.grad.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ga.cpu().numpy(), aud.grad.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gs.cpu().numpy(), sub.grad.numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("rows,dim", [(1, 8), (64, 512), (513, 1024), (7, 100)])
def test_l2norm_and_bwd(rows, dim):
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(rows + dim)
    x = torch.randn(rows, dim, generator=gen)
    x[0] = 0  # clamped row: x / eps
    send = torch.zeros(rows, 2 * dim, dtype=torch.bfloat16, device="cuda")
    y, inv = ops.l2norm(x.cuda(), out16=send[:, dim:], want_inv=True)
    ref = spec.l2_normalize(x.numpy())
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(send[:, dim:].float().cpu().numpy(), y.bfloat16().float().cpu().numpy())
    assert send[:, :dim].abs().max().item() == 0
    xr = x[1:].clone().requires_grad_()
    go = torch.randn(rows - 1, dim, generator=gen)
    torch.nn.functional.normalize(xr, dim=-1).backward(go)
    if rows > 1:
        gx = ops.l2norm_bwd(go.cuda(), y[1:], inv[1:])
        np.testing.assert_allclose(gx.cpu().numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_pack_pair():
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(5)
    t = torch.randn(33, 72, generator=gen)
    c = torch.randn(33, 72, generator=gen)
    p = ops.pack_pair(t.cuda(), c.cuda()).cpu()
    assert torch.equal(p[:, :72], t.bfloat16()) and torch.equal(p[:, 72:], c.bfloat16())


@pytest.mark.parametrize("dtype,S,H", [(torch.float32, 6, 16), (torch.bfloat16, 583, 768), (torch.float16, 3, 8)])
def test_gather_rows_concat3(dtype, S, H):
    from vast_b200 import ops
    gen = torch.Generator().manual_seed(6)
    bs, n, L, rank = 5, 15, 7, 1
    cond_all = torch.randn(n, S, H, generator=gen).to(dtype)
    ids_all = torch.randint(0, 30522, (n, L), generator=gen)
    mask_all = (torch.rand(n, L, generator=gen) > 0.3).long()
    sl = slice(rank * bs, (rank + 1) * bs)
    neg_text = torch.randint(0, n, (bs,), generator=gen)
    neg_cond = torch.randint(0, n, (bs,), generator=gen)
    ids1, att1, cond3 = ops.gather_rows_concat3(ids_all[sl].cuda(), mask_all[sl].cuda(), ids_all.cuda(), mask_all.cuda(),
                                                cond_all[sl].cuda(), cond_all.cuda(), neg_text.cuda(), neg_cond.cuda())
    r_ids, r_att, r_cond = spec.gather_negatives(cond_all[sl].float().numpy(), cond_all.float().numpy(), ids_all[sl].numpy(),
                                                 mask_all[sl].numpy(), ids_all.numpy(), mask_all.numpy(),
                                                 neg_cond.numpy(), neg_text.numpy())
    assert np.array_equal(ids1.cpu().numpy(), r_ids) and np.array_equal(att1.cpu().numpy(), r_att)
    assert np.array_equal(cond3.float().cpu().numpy(), r_cond)


def test_gather_matches_reference_golden(golden):
    from vast_b200 import ops
    g = golden("omc_w1")
    ids1, att1, cond3 = ops.gather_rows_concat3(cu(g["input_ids"]), cu(g["attention_mask"]), cu(g["input_ids"]),
                                                cu(g["attention_mask"]), cu(g["cond"]), cu(g["cond"]),
                                                cu(g["neg_cond2t"]), cu(g["neg_t2cond"]))
    assert np.array_equal(ids1.cpu().numpy(), g["input_ids_1"])
    assert np.array_equal(att1.cpu().numpy(), g["attention_mask_1"])
    assert np.array_equal(cond3.cpu().numpy(), g["condition_feats_3"])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_peer_row_exchange_emulated_ranks(dtype):
    """SURVEY 8 f-1 through the C-ABI with the ranks emulated on one GPU (every rank's block is a local buffer, the
    'peer' pointers are their addresses): vast_gather_rows_concat3_peer == cat(cond, all_gather(cond)[neg], cond), and
    vast_pull_row_grads == the gradient all_gather_with_grad + index would return, summed in request order."""
    from vast_b200 import ops
    world, bs, S, H, L = 3, 40, 7, 24, 5
    g = torch.Generator().manual_seed(8)
    blocks = [torch.randn(bs, S, H, generator=g).to(dtype).cuda() for _ in range(world)]
    allc = torch.cat(blocks)
    ids_all = torch.randint(0, 1000, (world * bs, L), generator=g).cuda()
    mask_all = torch.randint(0, 2, (world * bs, L), generator=g).cuda()
    negs_c = [torch.randint(0, world * bs, (bs,), generator=g).cuda() for _ in range(world)]
    negs_c[1][:20] = 7                                  # one row requested 20 times by rank 1 (> the 16-entry request list)
    negs_t = [torch.randint(0, world * bs, (bs,), generator=g).cuda() for _ in range(world)]
    ptrs = [b.data_ptr() for b in blocks]
    for r in range(world):
        sl = slice(r * bs, (r + 1) * bs)
        ids1, att1, cond3 = ops.gather_rows_concat3_peer(ids_all[sl], mask_all[sl], ids_all, mask_all, blocks[r], ptrs, bs,
                                                         negs_t[r], negs_c[r])
        assert torch.equal(cond3, torch.cat([blocks[r], allc[negs_c[r]], blocks[r]]))
        assert torch.equal(ids1, torch.cat([ids_all[sl], ids_all[sl], ids_all[negs_t[r]]]))
        assert torch.equal(att1, torch.cat([mask_all[sl], mask_all[sl], mask_all[negs_t[r]]]))
    # backward: every rank's gradient block of its fetched rows; owners pull
    gblocks = [torch.randn(bs, S, H, generator=g).to(dtype).cuda() for _ in range(world)]
    bases = [torch.randn(bs, S, H, generator=g).to(dtype).cuda() for _ in range(world)]
    req = torch.cat(negs_c)
    gptrs = [b.data_ptr() for b in gblocks]
    gall = torch.cat(gblocks).float()
    for r in range(world):
        out = ops.pull_row_grads(req, gptrs, bs, r * bs, gblocks[r], base_grad=bases[r])
        want, want0 = bases[r].float().clone(), torch.zeros_like(bases[r], dtype=torch.float32)
        for e in range(world * bs):                      # ascending request order, fp32 accumulation
            j = int(req[e])
            if r * bs <= j < (r + 1) * bs:
                want[j - r * bs] += gall[e]
                want0[j - r * bs] += gall[e]
        assert torch.equal(out, want.to(dtype))
        assert torch.equal(ops.pull_row_grads(req, gptrs, bs, r * bs, gblocks[r]), want0.to(dtype))   # no base gradient
