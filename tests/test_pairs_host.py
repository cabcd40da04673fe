"""CPU: host logic of the pair loop - the batched form of symmetric_inference
(sparse_ga.py:571-592 with a batch dimension) against the pair-by-pair form on a small stand-in network that offers the
three methods the glue calls (`_encode_image_pairs`, `_decoder`, `_downstream_head`), and which pairs forward_mast3r
decides to compute (sparse_ga.py:529-562: one network call per unordered pair, the mirrored pair reuses the matches)."""
import numpy as np
import torch

from starst3r_b200 import reconstruct as rc


class TinyNet:
    """Per-sample arithmetic only, so a batch must reproduce the single-pair results bit for bit."""

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.mix = torch.randn(3, 3, generator=g)
        self.head = torch.randn(3, 24, generator=g)
        self.calls = 0

    def _encode_image_pairs(self, im1, im2, shape1, shape2):
        self.calls += 1
        f = lambda im: im.flatten(2).transpose(1, 2) @ self.mix              # noqa: E731   [B, HW, 3]
        return f(im1), f(im2), torch.zeros(im1.shape[0], 1), torch.zeros(im2.shape[0], 1)

    def _decoder(self, fa, pa, fb, pb):
        return [fa, fa + 0.5 * fb.mean(1, keepdim=True)], [fb, fb - fa.amax(1, keepdim=True)]

    def _downstream_head(self, i, toks, shape):
        B = toks[-1].shape[0]
        H, W = int(shape[0, 0]), int(shape[0, 1])
        t = toks[-1].reshape(B, H, W, 3)
        conf = 1 + t.square().sum(-1)
        return {"pts3d": t * float(i), "conf": conf, "desc": torch.nn.functional.normalize(t @ self.head, dim=-1),
                "desc_conf": conf * 0.5}


def _imgs(n, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"img": torch.randn(1, 3, H, W, generator=g), "true_shape": np.int32([[H, W]]), "idx": i, "instance": f"{i}.png"}
            for i in range(n)]


def test_symmetric_inference_batch_equals_pair_by_pair():
    net = TinyNet()
    imgs = _imgs(4, 6, 8)
    i1, i2 = [imgs[0], imgs[2], imgs[3]], [imgs[1], imgs[0], imgs[2]]
    batched = rc.symmetric_inference_batch(net, i1, i2, "cpu")
    assert net.calls == 1 and len(batched) == 3                      # one pass of the network for the three pairs
    for (a, b), got in zip(zip(i1, i2), batched):
        want = rc.symmetric_inference(net, a, b, "cpu")
        assert len(got) == 4
        for rg, rw in zip(got, want):                                # res11, res21, res22, res12
            assert set(rg) == set(rw)
            for k in rw:
                assert rg[k].shape == rw[k].shape and rg[k].shape[0] == 1
                assert torch.equal(rg[k], rw[k]), k


def test_symmetric_inference_batch_falls_back_on_mixed_sizes_and_model_entry_points():
    net = TinyNet()
    a, b = _imgs(2, 6, 8), _imgs(2, 4, 8, seed=1)
    out = rc.symmetric_inference_batch(net, [a[0], b[0]], [a[1], b[1]], "cpu")       # two image sizes: pair by pair
    assert net.calls == 2 and out[0][0]["pts3d"].shape == (1, 6, 8, 3) and out[1][0]["pts3d"].shape == (1, 4, 8, 3)

    class Own:
        def symmetric_inference(self, x, y):
            return ("pair", x["idx"], y["idx"])

    assert rc.symmetric_inference_batch(Own(), [a[0], a[1]], [a[1], a[0]], "cpu") == [("pair", 0, 1), ("pair", 1, 0)]

    class OwnBatch(Own):
        def symmetric_inference_batch(self, xs, ys):
            return [("batch", x["idx"], y["idx"]) for x, y in zip(xs, ys)]

    assert rc.symmetric_inference_batch(OwnBatch(), [a[0]], [a[1]], "cpu") == [("batch", 0, 1)]


def test_forward_mast3r_computes_each_unordered_pair_once(monkeypatch):
    """The planning of forward_mast3r: complete symmetrised pair list in, one computation per unordered pair in the
    reference's loop order, every ordered pair in the result, entries already in the memo are not recomputed."""
    imgs = _imgs(4, 6, 8)
    pairs = [(imgs[i], imgs[j]) for i in range(4) for j in range(i)] + [(imgs[j], imgs[i]) for i in range(4) for j in range(i)]
    computed = []

    def fake_compute(memo, model, todo, device, desc_conf, subsample):
        for x, y in todo:
            a, b = x["instance"], y["instance"]
            computed.append((a, b))
            memo["fwd"][a, b] = memo["fwd"][b, a] = (torch.zeros(1),) * 4
            memo["corres"][a, b] = ((1.0, 2.0, 1), (torch.zeros(1, 2), torch.ones(1, 2), torch.ones(1)))

    monkeypatch.setattr(rc, "_compute_pairs", fake_compute)
    rc.clear_cache()
    res, cache = rc.forward_mast3r(pairs, object(), cache_path="plan-test", device="cpu")
    assert computed == [(x["instance"], y["instance"]) for x, y in pairs[:6]]
    assert list(res) == [(x["instance"], y["instance"]) for x, y in pairs]
    memo = rc._memo(cache)
    xy1, xy2, _ = memo["corres"]["0.png", "1.png"][1]                 # mirrored entry: xy1 / xy2 swapped (:538-540)
    assert torch.equal(xy1, torch.ones(1, 2)) and torch.equal(xy2, torch.zeros(1, 2))
    del computed[:]
    rc.forward_mast3r(pairs, object(), cache_path="plan-test", device="cpu")
    assert computed == []                                             # everything is in the memo now
    rc.clear_cache()
