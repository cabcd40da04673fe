"""CPU (gloo, world size 2): the multi-GPU host logic — pair / view sharding, the single gradient all-reduce,
variable-length gathers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from starst3r_b200 import dist as sd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # pair sharding: disjoint, complete
        mine = sd.shard_pairs(8)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        flat = [p for part in everyone for p in part]
        assert sorted(flat) == sorted(sd.unordered_pairs(8)) and len(set(flat)) == 28
        assert abs(len(everyone[0]) - len(everyone[1])) <= 1
        # view sharding
        assert sd.shard_indices(8) == list(range(rank, 8, world))
        # gradient all-reduce: every rank ends with the sum
        g = torch.Generator().manual_seed(rank)
        N = 50
        grads = {"means": torch.randn(N, 3, generator=g), "quats": torch.randn(N, 4, generator=g),
                 "scales": torch.randn(N, 3, generator=g), "opacities": torch.randn(N, generator=g),
                 "sh": torch.randn(N, 4, 3, generator=g)}
        ref = {}
        for k in grads:
            parts = []
            for r in range(world):
                gr = torch.Generator().manual_seed(r)
                t = {"means": torch.randn(N, 3, generator=gr), "quats": torch.randn(N, 4, generator=gr),
                     "scales": torch.randn(N, 3, generator=gr), "opacities": torch.randn(N, generator=gr),
                     "sh": torch.randn(N, 4, 3, generator=gr)}
                parts.append(t[k])
            ref[k] = sum(parts)
        sd.allreduce_gradients(grads)
        for k in grads:
            assert torch.allclose(grads[k], ref[k], atol=1e-6), k
        # variable-length gather of correspondence lists
        t = torch.arange((rank + 1) * 3 * 2, dtype=torch.int64).reshape(-1, 2) + 100 * rank
        parts = sd.gather_varlen(t)
        assert [p.shape[0] for p in parts] == [3 * (r + 1) for r in range(world)]
        assert torch.equal(parts[rank], t)
        x = torch.full((4,), float(rank))
        sd.broadcast_tensors([x], src=0)
        assert x.eq(0).all()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_passthrough():
    assert sd.world() == (0, 1)
    assert sd.shard_indices(5) == [0, 1, 2, 3, 4]
    g = {k: torch.ones(2) for k in sd.GRAD_KEYS}
    assert sd.allreduce_gradients(g) is g


def _shard_worker(rank, world, port, q):
    """Sharded-view training on the CPU oracle: the sum over ranks of the per-shard gradients (each rank renders only
    its views; the regularisers scale with the number of views it holds) equals the gradient of the full problem,
    which is what the GPU exchange (dist.PeerGradExchange / allreduce_gradients) relies on."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import gs_oracle as go
        from starst3r_b200 import synth
        torch.set_num_threads(1)
        W, H, C, N = 32, 24, 4, 60
        sp = synth.random_splats(N, seed=2, scale_mode="rand")
        sp["scales"] = sp["scales"] * 10
        viewmats, Ks = synth.look_at_cameras(C, W, H)
        truth = torch.rand(C, H, W, 3, generator=torch.Generator().manual_seed(0))
        mine = sd.shard_indices(C)

        def grads_of(idx):
            p = {k: v.clone() for k, v in sp.items()}
            st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in p.items()}
            _, g, _, _ = go.train_step(p, st, truth[idx], viewmats[idx], Ks[idx], W, H, 1)
            return {"means": g["means"], "quats": g["quats"], "scales": g["scales"], "opacities": g["opacities"],
                    "sh": g["shN"][:, :4].contiguous()}
        local = grads_of(mine)
        sd.allreduce_gradients(local)
        if rank == 0:
            full = grads_of(list(range(C)))
            for k in full:
                scale = float(full[k].abs().max()) + 1e-12
                assert float((full[k] - local[k]).abs().max()) <= 1e-4 * scale, k
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_views_gradient_sum_equals_full_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _oracle_extract(descs, qonfs, device=None, subsample=8):
    """CPU stand-in for match.extract_correspondences (which only runs on CUDA): the oracle restatement."""
    from oracle import match_oracle as mo
    xy1, xy2, conf = mo.extract_correspondences([d.cpu().numpy() for d in descs], [q.cpu().numpy() for q in qonfs], subsample)
    return torch.from_numpy(xy1), torch.from_numpy(xy2), torch.from_numpy(conf)


def _pairs_worker(rank, world, port, q):
    """forward_mast3r under a process group: every rank computes only its share of the image pairs (inference +
    matching) and ends with the full, identical memo."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from starst3r_b200 import match, synth
        from starst3r_b200 import reconstruct as rc
        torch.set_num_threads(1)
        match.extract_correspondences = _oracle_extract          # test-only: the product path refuses CPU tensors
        n, W, H = 4, 48, 32
        net = synth.SyntheticMast3r(n, W, H, seed=0, device="cpu", arc_deg=90.0)
        calls = []
        orig = net.symmetric_inference
        net.symmetric_inference = lambda a, b, device=None: (calls.append((a["idx"], b["idx"])), orig(a, b))[1]
        imgs = rc.prepare_images_for_mast3r(net.images())
        names = [f"{i}.png" for i in range(n)]

        def run(tag, shard):
            rc.SHARD_PAIRS = shard
            calls.clear()
            pairs = rc.convert_dust3r_pairs_naming(names, rc.make_pairs(imgs))
            res, _ = rc.forward_mast3r(pairs, net, cache_path=tag, device="cpu")
            return res, rc._memo(tag), list(calls)
        res_s, memo_s, calls_s = run("sharded", True)
        res_f, memo_f, calls_f = run("full", False)
        assert len(calls_f) == 6 and len(calls_s) == 6 // world, (calls_s, calls_f)  # 6 unordered pairs, 1/G of them here
        assert res_s == res_f
        assert set(memo_s["fwd"]) == set(memo_f["fwd"]) and set(memo_s["corres"]) == set(memo_f["corres"])
        # by-image ownership: image a's own full-resolution map lives only on rank index(a) mod G; the cross maps are
        # held by everybody at the optimiser's resolution
        resident = 0
        for (a, b), (X, C, X2, C2) in memo_f["fwd"].items():
            Xs, Cs, X2s, C2s = memo_s["fwd"][a, b]
            if names.index(a) % world == rank:
                assert torch.equal(Xs, X) and torch.equal(Cs, C), (a, b)
                resident += 1
            else:
                assert Xs is None and Cs is None, (a, b)
            assert isinstance(X2s, rc._Sub) and torch.equal(X2s.t, X2[::8, ::8]) and torch.equal(C2s.t, C2[::8, ::8]), (a, b)
        assert resident == sum(n - 1 for i in range(n) if i % world == rank)         # my images x their ordered pairs
        for k in memo_f["corres"]:
            (s0, s1, sn), (a1, a2, ac) = memo_s["corres"][k]
            (f0, f1, fn), (b1, b2, bc) = memo_f["corres"][k]
            assert sn == fn and abs(s0 - f0) < 1e-6 and abs(s1 - f1) < 1e-4
            assert torch.equal(a1, b1) and torch.equal(a2, b2) and torch.equal(ac, bc)
            assert sn > 10
        everyone = [None] * world
        dist.all_gather_object(everyone, sorted(calls_s))
        assert sorted(p for part in everyone for p in part) == sorted(calls_f)       # disjoint and complete
        # canonical views: image i is computed by rank i mod G and broadcast (oracle stand-ins for the two kernels)
        from oracle import align_oracle as ao
        canon_calls = []

        def canon_cpu(pts, cfs, subsample, mode="avg-angle"):
            canon_calls.append(len(pts))
            return ao.canonical_view(pts, cfs, subsample)
        rc.canonical_view = canon_cpu
        rc.estimate_focal_knowing_depth = lambda pts3d, pp, mode, min_focal, max_focal: ao.estimate_focal_weiszfeld(
            pts3d[0], min_focal, max_focal).reshape(1)
        out = {}
        for tag, shard, res in (("sharded", True, res_s), ("full", False, res_f)):
            rc.SHARD_PAIRS = shard
            canon_calls.clear()
            out[tag] = rc.prepare_canonical_data(names, res, 8, cache_path=tag, device="cpu", mode="avg-angle")
            assert len(canon_calls) == (len(range(rank, n, world)) if shard else n), (tag, canon_calls)
        (_, ps_s, cv_s, _, p21_s), (_, ps_f, cv_f, _, p21_f) = out["sharded"], out["full"]
        assert torch.equal(ps_s, ps_f)
        for i1 in p21_f:
            for i2 in p21_f[i1]:
                assert torch.equal(p21_s[i1][i2][0], p21_f[i1][i2][0]) and torch.equal(p21_s[i1][i2][1], p21_f[i1][i2][1])
        for img in names:
            pp_s, hw_s, f_s, core_s, _, idx_s, off_s = cv_s[img]
            pp_f, hw_f, f_f, core_f, _, idx_f, off_f = cv_f[img]
            assert hw_s == hw_f and torch.equal(f_s, f_f) and torch.equal(core_s, core_f)
            for other in idx_f:
                assert torch.equal(idx_s[other], idx_f[other]) and torch.equal(off_s[other], off_f[other])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_forward_mast3r_shards_pairs_and_exchanges_results(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pairs_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=600) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def _align_worker(rank, world, port, q):
    """sparse_scene_optimizer_slam under a process group: rank 0 alone runs the optimiser, every rank ends with rank 0's
    parameters and results (the kernels' fp32 atomics make independent replicas drift apart along the gauge)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from starst3r_b200 import _lib
        from starst3r_b200 import reconstruct as rc
        torch.set_num_threads(1)
        fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "align_match3.pt"),
                        weights_only=False)
        inp = fx["inputs"]
        _lib.require_cuda_device = lambda device, what="": torch.device("cpu")     # test-only: host stand-ins below
        seen = []

        def fake_phase(t, meta, params, mode, train_mask, gamma, lr_base, niter, schedule, dust3r_w, lossd_gamma,
                       want_grad=False):
            """Stand-in for the CUDA optimiser: results and parameter updates that depend on the rank that ran it."""
            seen.append((mode, niter))
            N = meta["N"]
            if niter:
                params["trans"] += 1.0 + rank
                params["quats"] += 0.25 * (1 + rank)
            val = 1000.0 * rank + 10.0 * mode + niter
            aoff, coff = meta["aoff"], np.cumsum([0] + meta["n_core"])
            res = dict(intrinsics=torch.full((N, 3, 3), val), cam2w=torch.full((N, 4, 4), val + 1),
                       depthmaps=[torch.full((int(coff[i + 1] - coff[i]),), val + 2) for i in range(N)],
                       pts3d=[torch.full((int(aoff[i + 1] - aoff[i]), 3), val + 3) for i in range(N)])
            return res, torch.zeros(max(niter, 1))[:niter], None
        rc._optimize_phase = fake_phase
        out = {}
        for shard in (True, False):
            rc.SHARD_PAIRS = shard
            seen.clear()
            _, res_c, res_f, params_ret = rc.sparse_scene_optimizer_slam(
                list(inp["imgs"]), 8, inp["imsizes"], inp["pps"].clone(), inp["base_focals"].clone(),
                [c.clone() for c in inp["core_depth"]], inp["anchors"], inp["corres"], inp["corres2d"], inp["preds_21"],
                None, inp["mst"], lr1=0.07, niter1=30, lr2=0.014, niter2=20, device="cpu", opt_depth=False,
                shared_intrinsics=False, matching_conf_thr=5.0, verbose=False)
            out[shard] = (list(seen), res_c, res_f, params_ret)
        seen_s, res_c, res_f, pr = out[True]
        assert seen_s == ([(0, 30), (1, 20)] if rank == 0 else [(0, 0), (1, 0)]), seen_s      # only rank 0 iterates
        # every rank holds what rank 0 computed: parameters after both phases, results of both phases
        assert all(torch.equal(x, torch.full_like(x, 2.0)) for x in pr["trans"])
        assert all(torch.allclose(x, torch.tensor([0.5, 0.5, 0.5, 1.5])) for x in pr["quats"])
        assert float(res_c["intrinsics"][0, 0, 0]) == 30.0 and float(res_c["cam2w"][0, 0, 0]) == 31.0
        assert float(res_f["intrinsics"][0, 0, 0]) == 30.0 and float(res_f["pts3d"][0][0, 0]) == 33.0
        assert all(float(d[0]) == 32.0 for d in res_f["depthmaps"])
        # without sharding (SHARD_PAIRS off) the call is the single-process one: every rank iterates for itself
        assert out[False][0] == [(0, 30), (1, 20)]
        assert float(out[False][2]["intrinsics"][0, 0, 0]) == 1000.0 * rank + 30.0
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def test_alignment_runs_on_rank0_and_is_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_align_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=600) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
