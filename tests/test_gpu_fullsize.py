"""-m gpu: BASELINE.json's FULL sizes, checked through size-independent properties (round trips, exact integer
checksums, additivity, idempotence) plus the oracle on a sample — the oracle alone would take minutes at these sizes.

configs[1]: 24 shards x 250 records of 256x256x3 u8 + 256x256 u8          (6000 records, 1.57 GB of shards)
configs[2]: 512x512x4 u16 + 512x512 u8 as float32-array records           (5.24 MB per record)
configs[4]: T=32 stacks of 256x256x4 u16 with date / cloud filters + fused band statistics
(configs[3], the 16x1024x1024x8 median, is in test_gpu_composite.py::test_median_full_size_properties.)
"""
import numpy as np
import pytest

from oracle import example_proto as oep
from oracle import normalise as onorm
from oracle import tfrecord as otfr

pytestmark = pytest.mark.gpu

H = W = 256
C, K = 3, 10
N_SHARDS, RECS = 24, 250


def _make_shard(torch, ops, dev, g, s):
    imgs = torch.randint(0, 256, (RECS, H, W, C), dtype=torch.uint8, device=dev, generator=g)
    labs = torch.randint(0, K, (RECS, H, W), dtype=torch.uint8, device=dev, generator=g)
    labs[torch.rand((RECS, H, W), device=dev, generator=g) < 0.02] = 255
    items = [dict(img=imgs[i].reshape(-1), tgt=labs[i].reshape(-1), kind=1, h=H, w=W, c=C, th=H, tw=W,
                  identifier=("256#2#1.0#43#%03d#%03d" % (s, i)).encode()) for i in range(RECS)]
    buf, offs, total = ops.build_records(items, dev)
    return buf[:total].clone(), imgs, labs, offs


def test_cfg2_full_size_round_trip_and_properties(dev):
    import torch

    from dl_image_segmentation_b200 import ops
    g = torch.Generator(device=dev)
    g.manual_seed(2002)
    mean = np.array([127.4, 126.9, 128.2], np.float32)
    std = np.array([73.9, 74.1, 73.6], np.float32)
    mean_d, std_d = torch.from_numpy(mean).to(dev), torch.from_numpy(std).to(dev)
    classes = torch.arange(K, device=dev, dtype=torch.uint8)
    acc = torch.zeros((C, 4), dtype=torch.int64, device=dev)
    want_n = want_s = want_ss = 0
    rec_bytes = None
    total_records = 0
    for s in range(N_SHARDS):
        shard, imgs, labs, offs = _make_shard(torch, ops, dev, g, s)
        if rec_bytes is None:
            rec_bytes = offs[1] - offs[0]
            # frame + Example of the first record: byte-exact against the oracle's serialiser and framer
            want = otfr.frame(oep.convert_to_example(imgs[0].cpu().numpy(), labs[0].cpu().numpy(), H, W, C, H, W,
                                                     "256#2#1.0#43#000#000").SerializeToString())
            assert bytes(shard[:rec_bytes].cpu().numpy()) == want
            assert rec_bytes == H * W * C + H * W + 231 + 20        # 262 397 B with its 22-char keys
        assert shard.numel() == RECS * rec_bytes
        st = ops.open_shard_async(shard, dev, max_records=RECS)
        # round trip: payload bytes as stored == what went in
        ib, tb, status = ops.parse_table(st, "raw", H * W * C, H * W, verify_crc=True)
        assert st.check("cfg2 shard") == RECS
        total_records += RECS
        assert not status.any()
        assert torch.equal(ib[:, :H * W * C].view(RECS, H, W, C), imgs)
        assert torch.equal(tb[:, :H * W].view(RECS, H, W), labs)
        # fused verify + normalise + one-hot == fp32 reference of the same op, bit for bit
        fi, ft, status = ops.parse_table(st, "norm_onehot", H * W * C, H * W, verify_crc=True, mean=mean_d, std=std_d,
                                         num_classes=K)
        assert not status.any()
        ref_i = (imgs.to(torch.float32) - mean_d) / std_d
        got_i = fi.view(RECS, H, W, C)
        assert torch.equal(got_i, ref_i)
        torch.testing.assert_close(got_i, ref_i, rtol=1e-6, atol=0)      # the north-star tolerance, stated
        got_t = ft.view(RECS, H, W, K)
        assert torch.equal(got_t, (labs.unsqueeze(-1) == classes).to(torch.float32))
        assert torch.equal(got_t.sum(-1) == 0, labs >= K)                   # nodata -> all-zero row
        if s in (0, 13):                                                    # and the NumPy oracle on a sample
            r = 7 * (s + 1)
            np.testing.assert_array_equal(got_i[r].cpu().numpy(), onorm.normalise(imgs[r].cpu().numpy()[None], mean, std)[0])
            np.testing.assert_array_equal(got_t[r].cpu().numpy(), onorm.one_hot(labs[r].cpu().numpy()[None], K)[0])
        # exact integer band statistics accumulate across shards
        ops.band_stats(imgs.view(-1, C), acc=acc, device=dev)
        x = imgs.view(-1, C).to(torch.int64)
        want_n += x.shape[0]
        want_s = want_s + x.sum(0)
        want_ss = want_ss + (x * x).sum(0)
        if s == 5:
            # one flipped payload bit in one of 250 records: exactly that record is reported, the rest still parse
            bad = shard.clone()
            bad[123 * rec_bytes + 4242] ^= 0x10
            st2 = ops.open_shard_async(bad, dev, max_records=RECS)
            _, _, status2 = ops.parse_table(st2, "norm_onehot", H * W * C, H * W, verify_crc=True, mean=mean_d, std=std_d,
                                            num_classes=K)
            assert torch.nonzero(status2).flatten().tolist() == [123]
            with pytest.raises(ops.DataLossError):
                st2.check("corrupted shard")
        del shard, imgs, labs, ib, tb, fi, ft
    assert total_records == 6000
    got = ops.stats_to_python(acc)
    for b in range(C):
        assert got[b] == (want_n, int(want_s[b]), int(want_ss[b]))


def test_cfg3_full_size_float_records_round_trip(dev):
    import torch

    from dl_image_segmentation_b200 import ops
    g = torch.Generator(device=dev)
    g.manual_seed(3003)
    n, S, B = 16, 512, 4
    imgs = torch.randint(0, 10047, (n, S, S, B), dtype=torch.int32, device=dev, generator=g).to(torch.int16).view(torch.uint16)
    labs = torch.randint(0, K, (n, S, S), dtype=torch.uint8, device=dev, generator=g)
    labs[torch.rand((n, S, S), device=dev, generator=g) < 0.02] = 255
    keys = ["448#32#10.0#43#%d#%d" % (7, i) for i in range(n)]
    items = [dict(img=imgs[i].reshape(-1), tgt=labs[i].reshape(-1), kind=2, h=S, w=S, c=B, th=S, tw=S,
                  identifier=keys[i].encode()) for i in range(n)]
    buf, offs, total = ops.build_records(items, dev)
    shard = buf[:total]
    # the first record, byte for byte, against the oracle (float32 FloatList of arr.flatten(), TFRecord frame)
    img0 = imgs[0].view(torch.int16).cpu().numpy().view(np.uint16)
    want = otfr.frame(oep.convert_to_example(img0, labs[0].cpu().numpy(), S, S, B, S, S, keys[0]).SerializeToString())
    assert len(want) == offs[1] - offs[0] and bytes(shard[:len(want)].cpu().numpy()) == want
    assert len(want) >= 4 * (S * S * B + S * S) and len(want) - 4 * (S * S * B + S * S) < 300
    st = ops.open_shard_async(shard, dev, max_records=n)
    ib, tb, status = ops.parse_table(st, "raw", S * S * B * 4, S * S * 4, verify_crc=True)
    assert st.check("cfg3 shard") == n and not status.any()
    got_i = ib[:, :S * S * B * 4].contiguous().view(torch.float32).view(n, S, S, B)
    got_t = tb[:, :S * S * 4].contiguous().view(torch.float32).view(n, S, S)
    assert torch.equal(got_i, imgs.view(torch.int16).to(torch.int32).to(torch.float32))
    assert torch.equal(got_t, labs.to(torch.float32))


def test_cfg5_mosaic_full_shape_properties(dev):
    import torch

    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(5005)
    n, T, S, B = 48, 32, 256, 4
    g = torch.Generator(device=dev)
    g.manual_seed(5005)
    stacks = torch.randint(0, 65536, (n, T, S, S, B), dtype=torch.int32, device=dev, generator=g).to(torch.uint16)
    valids = (torch.rand((n, T, S, S), device=dev, generator=g) < 0.85).to(torch.uint8)
    days = np.sort(rng.integers(0, 730, (n, T)), axis=1).astype(np.int32)
    cfs = rng.random((n, T)).astype(np.float32)
    cfs[3] = 0.9                                               # a chip whose scenes are all too cloudy
    flt = dict(ref_day=365, min_day=180, max_day=545, max_cf=0.4)
    acc = torch.zeros((B, 4), dtype=torch.int64, device=dev)
    out, mask, src, nel = ops.nearest_date_mosaic(stacks, valids, days, cfs, device=dev, stats_acc=acc, **flt)
    elig = (days >= 180) & (days < 545) & (cfs < np.float32(0.4))          # start inclusive, end exclusive, strict <
    assert nel.cpu().numpy().tolist() == elig.sum(1).tolist() and nel[3] == 0
    elig_d = torch.from_numpy(elig).to(dev)
    ok = valids.bool() & elig_d[:, :, None, None]                           # (n,T,S,S): scene may paint this pixel
    assert torch.equal(mask, ~ok.any(1))
    # the chosen scene is eligible + valid there, and the output is that scene's pixel
    s64 = src.to(torch.int64).clamp(min=0)
    assert torch.equal(torch.gather(ok, 1, s64[:, None]).squeeze(1), ~mask)
    picked = torch.gather(stacks.view(torch.int16), 1, s64[:, None, :, :, None].expand(n, 1, S, S, B)).squeeze(1)
    assert torch.equal(torch.where(mask[..., None], torch.zeros_like(picked), picked), out.view(torch.int16))
    # ... and no other paintable scene is closer to the reference date; ties go to the later scene
    dist = torch.from_numpy(np.abs(days - 365)).to(dev)[:, :, None, None].expand(n, T, S, S)
    big = torch.full_like(dist, 1 << 30)
    best = torch.where(ok, dist, big).min(1).values
    assert torch.equal(torch.gather(dist, 1, s64[:, None]).squeeze(1)[~mask], best[~mask])
    t_idx = torch.arange(T, device=dev)[None, :, None, None].expand(n, T, S, S)
    last_best = torch.where(ok & (dist == best[:, None]), t_idx, torch.full_like(t_idx, -1)).max(1).values
    assert torch.equal(s64[~mask], last_best[~mask])
    # the oracle's painter's loop on two whole chips
    from oracle import composite as ocomp
    for i in (0, 17):
        r_out, r_mask, r_src = ocomp.nearest_date_mosaic(stacks[i].view(torch.int16).cpu().numpy().view(np.uint16),
                                                         valids[i].cpu().numpy(), days[i], cfs[i], 365, 180, 545, 0.4)
        np.testing.assert_array_equal(out[i].view(torch.int16).cpu().numpy().view(np.uint16), r_out)
        np.testing.assert_array_equal(mask[i].cpu().numpy(), r_mask)
    # fused statistics: exact integers, and additive over any split of the chips (what the one allreduce relies on)
    o64 = out.view(torch.int16).to(torch.int64) & 0xFFFF
    keep = (~mask)[..., None].to(torch.int64)
    want = [(int((~mask).sum()), int((o64[..., b] * keep[..., 0]).sum()), int((o64[..., b] ** 2 * keep[..., 0]).sum()))
            for b in range(B)]
    assert ops.stats_to_python(acc) == want
    acc2 = torch.zeros_like(acc)
    for lo, hi in ((0, 20), (20, 48)):
        ops.nearest_date_mosaic(stacks[lo:hi], valids[lo:hi], days[lo:hi], cfs[lo:hi], device=dev, stats_acc=acc2,
                                want_src=False, **flt)
    assert ops.stats_to_python(acc2) == want             # (the two sum-of-squares counters are not canonical on their own)
    # idempotence: a mosaic of the mosaic (one always-eligible scene, validity = not masked) is the mosaic
    out2, mask2, _, _ = ops.nearest_date_mosaic(out[:, None].contiguous(), (~mask).to(torch.uint8)[:, None].contiguous(),
                                                np.full((n, 1), 365, np.int32), np.zeros((n, 1), np.float32),
                                                device=dev, **flt)
    assert torch.equal(mask2, mask) and torch.equal(out2, out)
