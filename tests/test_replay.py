"""the in-HBM replay store (nextbestpath_b200/replay.py) against the reference's LMDB functions
(next_best_path/utility/nbp_utils.py:32-141) and experience assembly (:653-683, :741-756).  The store is a data structure (torch
tensors in device memory, no kernel of its own), so its logic is tested on the CPU device here; the GPU test feeds a training
micro-batch from it.  The reference functions are exec'd from /root/reference (build container) over a minimal in-memory stand-in
for the LMDB environment."""
import math
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from nextbestpath_b200.replay import PathExperiences, ReplayRing, micro_batch_loss

HERE = os.path.dirname(os.path.abspath(__file__))
HAVE_REF = os.path.exists("/root/reference/next_best_path/utility/nbp_utils.py")
S = 16


def _record(i, k=None, g=None):
    g = g or torch.Generator().manual_seed(1000 + i)
    k = (i % 5) + 1 if k is None else k
    return {"current_model_input": torch.randint(0, 40, (1, 5, S, S), generator=g).float(),
            "current_gt_2d_layout": (torch.rand(1, 1, S, S, generator=g) < 0.3).float(),
            "target_value_map_pixel": torch.stack((torch.randint(0, 8, (k,), generator=g), torch.randint(0, S // 4, (k,), generator=g),
                                                   torch.randint(0, S // 4, (k,), generator=g)), dim=-1),
            "actual_coverage_gain": torch.rand(k, generator=g) * 10, "pose_i": i}


def test_ring_store_batch_records_and_wraparound():
    ring = ReplayRing(6, "cpu", S=S, max_targets=8)
    recs = [_record(i) for i in range(9)]
    for r in recs:
        ring.store_experience(r)
    assert len(ring) == 6 and ring.next_key == 9                     # 3 oldest overwritten
    live = recs[3:]
    out = ring.records(range(6))
    for a, b in zip(out, live):
        assert np.array_equal(a["current_model_input"], b["current_model_input"].numpy()) and a["current_model_input"].shape == (1, 5, S, S)
        assert np.array_equal(a["current_gt_2d_layout"], b["current_gt_2d_layout"].numpy())
        assert np.array_equal(a["target_value_map_pixel"], b["target_value_map_pixel"].numpy()) and a["target_value_map_pixel"].dtype == np.int64
        assert np.array_equal(a["actual_coverage_gain"], b["actual_coverage_gain"].numpy()) and int(a["pose_i"]) == b["pose_i"]
    # batch() == what train_experience_data assembles from the records (nbp_utils.py:366-376)
    idx = [4, 0, 5]
    bt = ring.batch(idx)
    sel = [live[i] for i in idx]
    assert torch.equal(bt["inputs"], torch.cat([r["current_model_input"] for r in sel]))
    assert torch.equal(bt["layouts"], torch.cat([r["current_gt_2d_layout"] for r in sel]))
    assert torch.equal(bt["coords"], torch.cat([r["target_value_map_pixel"] for r in sel]))
    assert torch.equal(bt["gains"], torch.cat([r["actual_coverage_gain"] for r in sel]))
    sizes = [len(r["target_value_map_pixel"]) for r in sel]
    assert torch.equal(bt["sample_of"], torch.repeat_interleave(torch.arange(3), torch.tensor(sizes)))
    with pytest.raises(IndexError):
        ring.batch([6])
    # checkpoint / resume
    ring2 = ReplayRing(10, "cpu", S=S, max_targets=8)
    ring2.load_state_dict(ring.state_dict())
    assert len(ring2) == 6 and ring2.next_key == 9
    assert all(np.array_equal(a["current_model_input"], b["current_model_input"]) for a, b in zip(ring2.records(range(6)), out))
    # int16 store format
    r16 = ReplayRing(3, "cpu", S=S, max_targets=8, input_dtype=torch.int16)
    r16.store_experience(recs[0])
    assert torch.equal(r16.batch([0])["inputs"], recs[0]["current_model_input"])
    big = dict(recs[1]); big["current_model_input"] = recs[1]["current_model_input"] + 40000
    with pytest.raises(RuntimeError):
        r16.store_experience(big)


# ---------------------------------------------------------------------------------------------------- the reference's LMDB functions
class _Cursor:
    def __init__(self, db):
        self.db, self.pos = db, -1

    def _keys(self):
        return sorted(self.db)

    def __iter__(self):
        for k in self._keys():
            yield k, self.db[k]

    def first(self):
        self.pos = 0
        return bool(self.db)

    def last(self):
        self.pos = len(self.db) - 1
        return bool(self.db)

    def prev(self):
        if self.pos <= 0:
            return False
        self.pos -= 1
        return True

    def value(self):
        return self.db[self._keys()[self.pos]]


class _Txn:
    def __init__(self, db):
        self.db = db

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def cursor(self):
        return _Cursor(self.db)

    def put(self, k, v):
        self.db[k] = v

    def delete(self, k):
        del self.db[k]


class _Env:
    def __init__(self):
        self.db = {}

    def begin(self, write=False):
        return _Txn(self.db)

    def stat(self):
        return {"entries": len(self.db)}


def _reference_store_functions():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden as mg
    clock = {"t": 1.0}

    def fake_time():
        clock["t"] += 0.001                                          # one record per "millisecond": no key collisions
        return clock["t"]

    ns = {"time": types.SimpleNamespace(time=fake_time), "np": np, "random": random, "math": math,
          "msgpack": types.SimpleNamespace(packb=lambda d, use_bin_type=True: d, unpackb=lambda v, object_hook=None: v),
          "m": types.SimpleNamespace(decode=None)}
    exec(mg._ref_lines("/root/reference/next_best_path/utility/nbp_utils.py", "def store_experience(env, data):", "return combined_data"), ns)
    return ns


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container only)")
def test_selection_rules_equal_the_reference_lmdb_functions():
    ns = _reference_store_functions()
    env = _Env()
    ring = ReplayRing(64, "cpu", S=S, max_targets=8)
    for i in range(37):
        r = _record(i)
        ns["store_experience"](env, r)
        ring.store_experience(r)
    ids = lambda recs: [int(r["pose_i"]) for r in recs]
    # the stored record is the reference's (numpy, same shapes / dtypes)
    first = ns["read_combined_data"](env, sample_m=None)[0]
    mine = ring.records([0])[0]
    assert all(np.array_equal(first[k], mine[k]) and first[k].dtype == mine[k].dtype for k in first if k != "pose_i")
    for sample_m, sample_size in ((10, 2176 * 2), (50, 2176 * 2), (None, 0)):
        random.seed(5)
        want = ids(ns["read_combined_data"](env, sample_m=sample_m))
        random.seed(5)
        got = ring.read_combined_data(sample_m=sample_m)
        assert want == [int(ring.records([i])[0]["pose_i"]) for i in got], sample_m
    random.seed(7)
    want = ids(ns["read_random_data_readonly"](env, num_samples=9))
    random.seed(7)
    assert want == ring.read_random(9)
    assert ids(ns["store_validation_data_readonly"](env, num=8)) == ring.validation_indices(8)
    want = ids(ns["store_validation_data"](env, num=8))             # removes them from the database
    got = ring.take_validation(8)
    assert want == ids(got) and len(ring) == len(env.db) == 37 - len(want)
    assert ids(ns["read_combined_data"](env, sample_m=None)) == ids(ring.records(range(len(ring))))


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container only)")
def test_path_experience_flush_equals_reference_lines():
    """nbp_utils.py:653-683 exec'd from the reference tree with the reference's own map functions, against PathExperiences."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden as mg
    _, ru = mg.import_reference()
    g = torch.Generator().manual_seed(3)
    n = 9
    poses = torch.zeros(n, 5)
    poses[:, 0] = torch.cumsum(torch.randn(n, generator=g) * 9, 0)
    poses[:, 2] = torch.cumsum(torch.randn(n, generator=g) * 9, 0)
    poses[:, 1] = 1.8
    cov = torch.cumsum(torch.rand(n, generator=g) * 0.02 - 0.004, 0).tolist()          # some negative increments
    heads = torch.randint(0, 8, (n,), generator=g)
    inputs = [torch.randint(0, 30, (1, 5, S, S), generator=g).float() for _ in range(n)]
    gts = [(torch.rand(1, 1, S, S, generator=g) < 0.3).float() for _ in range(n)]
    replans = {0, 4}
    stored_ref, cur_in, exp_list = [], None, []
    for i in range(n):                                                # the two append sites (:741-747 / :749-755), stale input included
        if i in replans:
            cur_in = inputs[i]
        exp_list.append([cov[i], cur_in, gts[i], poses[i], heads[i]])
    ns = {"torch": torch, "experiences_list": exp_list, "device": "cpu", "value_map_size": (64, 64), "prediction_range": (-40, 40), "pose_i": 23,
          "transform_points_to_n_pieces": ru.transform_points_to_n_pieces, "get_point_position_in_the_img": ru.get_point_position_in_the_img,
          "store_experience": lambda env, d: stored_ref.append(d), "db_env": None}
    exec(mg._ref_lines("/root/reference/next_best_path/utility/nbp_utils.py", "if len(experiences_list) > 0:", "experiences_list = []"), ns)
    stored = []
    pe = PathExperiences(transform=ru.transform_points_to_n_pieces, cell_of=ru.get_point_position_in_the_img)
    for i in range(n):
        if i in replans:
            pe.on_replan(cov[i], inputs[i], gts[i], poses[i], heads[i])
        else:
            pe.on_path_step(cov[i], gts[i], poses[i], heads[i], model_input=inputs[i])
    assert pe.flush(stored.append, 23) == len(stored_ref) > 3 and len(pe) == 0
    for a, b in zip(stored, stored_ref):
        assert a["pose_i"] == b["pose_i"]
        for k in ("current_model_input", "current_gt_2d_layout", "target_value_map_pixel", "actual_coverage_gain"):
            assert torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype, k
    assert torch.equal(stored[2]["current_model_input"], inputs[0])   # the quirk: a mid-path entry stores the re-plan's input
    fixed = []
    pe = PathExperiences(stale_input=False, transform=ru.transform_points_to_n_pieces, cell_of=ru.get_point_position_in_the_img)
    for i in range(n):
        (pe.on_replan(cov[i], inputs[i], gts[i], poses[i], heads[i]) if i in replans
         else pe.on_path_step(cov[i], gts[i], poses[i], heads[i], model_input=inputs[i]))
    pe.flush(fixed.append, 23)
    assert torch.equal(fixed[2]["current_model_input"], inputs[2])


@pytest.mark.gpu
def test_training_micro_batch_from_the_hbm_store():
    """collect -> store -> train without leaving HBM: PathExperiences with the CUDA map shims fills a ring on the device; a micro-batch
    from ``ReplayRing.batch`` gives the same loss and gradients as the reference-layout records fed through the reference's
    micro-batch assembly (oracle.driver_lines.train_experience_data restates it; pinned in test_dropin_lines.py)."""
    from nextbestpath_b200.networks import NBP
    from oracle import driver_lines as DL
    from oracle import nbp_torch as NT
    dev = "cuda:0"
    Sg = 64
    ring = ReplayRing(32, dev, S=Sg, max_targets=16)
    g = torch.Generator().manual_seed(11)
    pe = PathExperiences(value_map_size=(Sg // 4, Sg // 4))
    for path in range(2):
        n = 6
        poses = torch.zeros(n, 5)
        poses[:, 0] = torch.cumsum(torch.randn(n, generator=g) * 6, 0); poses[:, 2] = torch.cumsum(torch.randn(n, generator=g) * 6, 0)
        cov = torch.cumsum(torch.rand(n, generator=g) * 0.03, 0).tolist()
        for i in range(n):
            x = NT.count_like_input(1, Sg, seed=50 + 10 * path + i).to(dev)
            gt = (torch.rand(1, 1, Sg, Sg, generator=g) < 0.2).float().to(dev)
            hd = torch.randint(0, 8, (1,), generator=g)[0].to(dev)
            (pe.on_replan if i == 0 else pe.on_path_step)(*((cov[i], x, gt, poses[i].to(dev), hd) if i == 0 else (cov[i], gt, poses[i].to(dev), hd)))
        assert pe.flush(ring, pose_i=20 + path) > 0
    assert len(ring) >= 6 and ring.inputs.is_cuda
    idx = [1, 4, 0, 5]
    grads = []
    losses = []
    for how in ("ring", "records"):
        net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(dev).train()
        if how == "ring":
            loss = micro_batch_loss(net, ring.batch(idx))
            loss.backward()
            losses.append(float(loss))
        else:
            opt = DL.RecordingAdamW(net.parameters(), lr=1e-3)
            random.seed(1)
            recs = ring.records(idx)
            order = list(range(len(recs))); random.shuffle(order); random.seed(1)      # train_experience_data shuffles: undo it
            inv = [recs[order.index(i)] for i in range(len(recs))]
            losses.append(DL.train_experience_data(inv, types.SimpleNamespace(nbp_batch_size=len(recs)), opt, net, dev, 2, torch.cuda.amp.GradScaler)[0] * 8)
            grads.append([gr.clone() for gr in opt.recorded])
            continue
        grads.append([p.grad.clone() for p in net.parameters()])
    assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[1])
    num = sum(float((a - b).double().pow(2).sum()) for a, b in zip(*grads)) ** 0.5
    den = sum(float(b.double().pow(2).sum()) for b in grads[1]) ** 0.5
    assert num <= 1e-3 * den, (num, den)            # the sample ORDER inside the micro-batch may differ: BatchNorm statistics do not depend on it
