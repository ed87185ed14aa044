"""In-HBM replay store for NBP training data.

The reference keeps its experiences in LMDB: every record is msgpack-serialised from device tensors to numpy to disk
(``store_experience`` next_best_path/utility/nbp_utils.py:32-42), read back record by record (``read_combined_data`` :99-141,
``read_random_data_readonly`` :59-71, ``store_validation_data`` :73-97) and converted numpy -> tensor -> device again for every
micro-batch (``train_experience_data`` :352-376).  Here the records live in preallocated device tensors (a ring: the oldest
record is overwritten when the store is full), the read functions return *indices* with the reference's selection rules, and
``batch(indices)`` assembles a micro-batch with a handful of device gathers -- nothing leaves HBM between collection and training
(BASELINE configs[2]: 512 tiles per optimizer step).

Record layout = the reference's (``experience_db`` :676-682):
    current_model_input (1,5,S,S) fp32 counts | current_gt_2d_layout (1,1,S,S) 0/1 | target_value_map_pixel (K,3) long
    (heading channel, gx, gy) | actual_coverage_gain (K,) fp32 | pose_i
Keys: the reference uses the wall-clock millisecond as the LMDB key, i.e. insertion order; here a monotonically increasing counter.

``PathExperiences`` is the collection side: the per-path experience list of ``trajectory_collection`` (:741-756) and its
"every later pose of the path is a target of every earlier one" augmentation (:653-683).
"""
from __future__ import annotations

import math
import random as _random

import numpy as np
import torch


class ReplayRing:
    def __init__(self, capacity: int, device, S: int = 256, n_channels: int = 5, max_targets: int = 128, input_dtype=torch.float32):
        """``input_dtype``: torch.float32 keeps the count images as the reference stores them (1.31 MB per record at S = 256:
        10 000 records = 13 GB of the 180 GB); torch.int16 halves that for counts < 32768 (checked on store)."""
        if input_dtype not in (torch.float32, torch.int16):
            raise ValueError("input_dtype must be torch.float32 or torch.int16")
        self.dev = torch.device(device)
        self.capacity, self.S, self.C, self.K = int(capacity), int(S), int(n_channels), int(max_targets)
        dev = self.dev
        self.inputs = torch.zeros((capacity, n_channels, S, S), dtype=input_dtype, device=dev)
        self.layouts = torch.zeros((capacity, 1, S, S), dtype=torch.uint8, device=dev)
        self.targets = torch.zeros((capacity, max_targets, 3), dtype=torch.int16, device=dev)
        self.gains = torch.zeros((capacity, max_targets), dtype=torch.float32, device=dev)
        self.n_targets = torch.zeros(capacity, dtype=torch.int32, device=dev)
        self.pose_i = torch.zeros(capacity, dtype=torch.int32, device=dev)
        self.keys = torch.zeros(capacity, dtype=torch.int64, device=dev)
        self._n_targets_host = [0] * capacity          # host mirror: batch() sizes its outputs without a device sync
        self._pose_i_host = [0] * capacity
        self.start, self.count, self.next_key = 0, 0, 0
        self.dropped_targets = 0                       # targets beyond max_targets (never hit with the reference's path lengths)

    # ------------------------------------------------------------------ bookkeeping
    def __len__(self):
        return self.count

    def _phys(self, logical):
        """logical index (0 = oldest) -> slot; accepts ints, lists, tensors."""
        if isinstance(logical, int):
            if not (0 <= logical < self.count):
                raise IndexError(logical)
            return (self.start + logical) % self.capacity
        idx = torch.as_tensor(logical, dtype=torch.int64)
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self.count):
            raise IndexError("replay index out of range")
        return (idx + self.start) % self.capacity

    # ------------------------------------------------------------------ write side
    def store_experience(self, data: dict) -> int:
        """store_experience(env, data) (nbp_utils.py:32-42) without the serialisation: ``data`` holds the reference's five fields as
        tensors (any device).  Returns the record's key."""
        x = torch.as_tensor(data["current_model_input"])
        lay = torch.as_tensor(data["current_gt_2d_layout"])
        pix = torch.as_tensor(data["target_value_map_pixel"]).reshape(-1, 3)
        gain = torch.as_tensor(data["actual_coverage_gain"]).reshape(-1)
        if tuple(x.shape[-3:]) != (self.C, self.S, self.S) or tuple(lay.shape[-2:]) != (self.S, self.S):
            raise RuntimeError(f"record shapes {tuple(x.shape)} / {tuple(lay.shape)} do not match the store ({self.C},{self.S},{self.S})")
        if pix.shape[0] != gain.shape[0]:
            raise RuntimeError("target_value_map_pixel and actual_coverage_gain differ in length")
        k = int(pix.shape[0])
        if k > self.K:
            self.dropped_targets += k - self.K
            pix, gain, k = pix[: self.K], gain[: self.K], self.K
        if self.count == self.capacity:                 # full: overwrite the oldest
            slot = self.start
            self.start = (self.start + 1) % self.capacity
        else:
            slot = (self.start + self.count) % self.capacity
            self.count += 1
        x = x.reshape(self.C, self.S, self.S).to(self.dev)
        if self.inputs.dtype == torch.int16:
            if float(x.max()) > 32767.0:
                raise RuntimeError("count image exceeds the int16 store format; use input_dtype=torch.float32")
            self.inputs[slot].copy_(x)
        else:
            self.inputs[slot].copy_(x.to(torch.float32))
        self.layouts[slot, 0].copy_(lay.reshape(self.S, self.S).to(self.dev) != 0)
        if k:
            self.targets[slot, :k].copy_(pix.to(self.dev))
            self.gains[slot, :k].copy_(gain.to(self.dev, torch.float32))
        self.n_targets[slot] = k
        pose_i = int(data["pose_i"])
        self.pose_i[slot] = pose_i
        self.keys[slot] = self.next_key
        self._n_targets_host[slot], self._pose_i_host[slot] = k, pose_i
        self.next_key += 1
        return self.next_key - 1

    def _delete(self, logical_sorted):
        """Remove records (ascending logical indices) and close the gaps, preserving order (LMDB ``txn.delete``)."""
        if not logical_sorted:
            return
        keep = [i for i in range(self.count) if i not in set(logical_sorted)]
        src = self._phys(keep).to(self.dev)
        for name in ("inputs", "layouts", "targets", "gains", "n_targets", "pose_i", "keys"):
            t = getattr(self, name)
            t[: len(keep)] = t[src].clone()
        nt = [self._n_targets_host[int(s)] for s in src.tolist()]
        pi = [self._pose_i_host[int(s)] for s in src.tolist()]
        self._n_targets_host[: len(keep)], self._pose_i_host[: len(keep)] = nt, pi
        self.start, self.count = 0, len(keep)

    # ------------------------------------------------------------------ read side: the reference's selection rules, as indices
    def read_combined_data(self, sample_m=2304 * 2, sample_size: int = 2176 * 2, rng=_random):
        """read_combined_data (nbp_utils.py:99-141): a random subset (``random.sample``, in store order) of the records older than
        the last ``sample_m``, followed by the last ``sample_m`` records.  ``sample_m=None``: everything."""
        total = self.count
        if sample_m is None:
            return list(range(total))
        n = total - sample_m
        if n < 0:
            n = 1
        chosen = set(rng.sample(range(n), min(sample_size, n)))
        selected = [i for i in range(min(n, total)) if i in chosen]
        last = list(range(max(total - sample_m, 0), total)) if total else []
        return selected + last

    def read_random(self, num_samples: int = 64, rng=_random):
        """read_random_data_readonly (nbp_utils.py:59-71): ``num_samples`` distinct records, in store order."""
        return sorted(rng.sample(range(self.count), num_samples))

    def validation_indices(self, num: int = 600 * 2):
        """store_validation_data_readonly (nbp_utils.py:44-57): every n-th record, n = ceil(total / num), at most ``num``."""
        total = self.count
        if total == 0:
            return []
        n = math.ceil(total / num)
        return [i for i in range(total) if i % n == 0][:num]

    def take_validation(self, num: int = 600 * 2):
        """store_validation_data (nbp_utils.py:73-97): the same selection, REMOVED from the store.  Returns the records
        (reference layout, numpy) since their slots are reused."""
        idx = self.validation_indices(num)
        recs = self.records(idx)
        self._delete(idx)
        return recs

    # ------------------------------------------------------------------ hand-off to training
    def batch(self, indices):
        """The tensors ``train_experience_data`` builds for one micro-batch (nbp_utils.py:366-376), straight from HBM:
        inputs (b,5,S,S) fp32, layouts (b,1,S,S) fp32, coords (T,3) long, gains (T,) fp32, sample_of (T,) long
        (``batch_indices``: which sample each target belongs to)."""
        slots = self._phys(list(indices))
        counts = [self._n_targets_host[int(s)] for s in slots.tolist()]
        sl = slots.to(self.dev)
        inputs = self.inputs[sl].to(torch.float32)
        layouts = self.layouts[sl].to(torch.float32)
        kk = torch.arange(self.K, device=self.dev).view(1, -1)
        live = kk < self.n_targets[sl].view(-1, 1)                                   # (b, K) mask of real targets, row-major = reference order
        coords = self.targets[sl][live].to(torch.int64)
        gains = self.gains[sl][live]
        sample_of = torch.repeat_interleave(torch.arange(len(counts), device=self.dev), torch.tensor(counts, device=self.dev))
        return {"inputs": inputs, "layouts": layouts, "coords": coords, "gains": gains, "sample_of": sample_of}

    def records(self, indices):
        """Records in the reference's msgpack layout (dicts of numpy arrays, what ``msgpack.unpackb(..., object_hook=m.decode)``
        yields): lets the reference's own ``train_experience_data`` consume the store unchanged."""
        out = []
        for i in indices:
            s = self._phys(int(i))
            k = self._n_targets_host[s]
            out.append({"current_model_input": self.inputs[s].to(torch.float32).unsqueeze(0).cpu().numpy(),
                        "current_gt_2d_layout": self.layouts[s].to(torch.float32).unsqueeze(0).cpu().numpy(),
                        "target_value_map_pixel": self.targets[s, :k].to(torch.int64).cpu().numpy(),
                        "actual_coverage_gain": self.gains[s, :k].cpu().numpy(),
                        "pose_i": np.array(self._pose_i_host[s])})
        return out

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        order = self._phys(list(range(self.count))).to(self.dev)
        return {"S": self.S, "C": self.C, "K": self.K, "capacity": self.capacity, "next_key": self.next_key,
                **{name: getattr(self, name)[order].cpu() for name in ("inputs", "layouts", "targets", "gains", "n_targets", "pose_i", "keys")}}

    def load_state_dict(self, sd):
        if (sd["S"], sd["C"], sd["K"]) != (self.S, self.C, self.K):
            raise RuntimeError("replay checkpoint geometry differs from this store")
        n = int(sd["keys"].shape[0])
        if n > self.capacity:
            raise RuntimeError(f"checkpoint holds {n} records, capacity is {self.capacity}")
        for name in ("inputs", "layouts", "targets", "gains", "n_targets", "pose_i", "keys"):
            getattr(self, name)[:n].copy_(sd[name].to(getattr(self, name).dtype))
        self.start, self.count, self.next_key = 0, n, int(sd["next_key"])
        self._n_targets_host[:n] = [int(v) for v in sd["n_targets"].tolist()]
        self._pose_i_host[:n] = [int(v) for v in sd["pose_i"].tolist()]


def micro_batch_loss(model, batch):
    """Forward + the sparse gather + NBP.loss of nbp_utils.py:378-381 on a ``ReplayRing.batch``."""
    value_map, obstacle_map = model(batch["inputs"])
    c = batch["coords"]
    picked = value_map[batch["sample_of"], c[:, 0], c[:, 1], c[:, 2]]
    return model.loss(picked, batch["gains"], obstacle_map, batch["layouts"])


class PathExperiences:
    """The experience list of one trajectory between two re-plans (``experiences_list``, nbp_utils.py:741-756) and its flush
    into the store (:653-683).

    Reference quirk, kept by default (``stale_input=True``): the entry appended at a re-plan carries the model input computed
    there (:693,:741-747), but the entries appended while the camera follows the path (:749-755) carry *the same, stale*
    ``current_model_input`` of the last re-plan together with the fresh obstacle map, pose and coverage.  ``stale_input=False``
    stores the input of the pose the entry belongs to instead (a deliberate fix; changes the training distribution)."""

    def __init__(self, value_map_size=(64, 64), prediction_range=(-40, 40), stale_input: bool = True, transform=None, cell_of=None):
        if transform is None or cell_of is None:
            from .utility import utils as U
            transform, cell_of = transform or U.transform_points_to_n_pieces, cell_of or U.get_point_position_in_the_img
        self.transform, self.cell_of = transform, cell_of
        self.value_map_size, self.prediction_range, self.stale_input = tuple(value_map_size), tuple(prediction_range), stale_input
        self.entries = []                 # [coverage, model_input, gt_obs, pose(5), heading index]
        self._replan_input = None

    def __len__(self):
        return len(self.entries)

    def on_replan(self, coverage: float, model_input, gt_obs, pose, heading_idx):
        """A new path was planned from this pose (:741-747)."""
        self._replan_input = model_input
        self.entries.append([float(coverage), model_input, gt_obs, pose, heading_idx])

    def on_path_step(self, coverage: float, gt_obs, pose, heading_idx, model_input=None):
        """The camera advanced along the current path (:749-755)."""
        if self.stale_input or model_input is None:
            if self._replan_input is None:
                raise RuntimeError("on_path_step before the first on_replan")
            model_input = self._replan_input
        self.entries.append([float(coverage), model_input, gt_obs, pose, heading_idx])

    def flush(self, store, pose_i: int) -> int:
        """End of the path (:653-685): entry ``e`` becomes a record whose targets are all LATER entries ``n`` of the path that fall
        inside e's value map: pixel (heading of n, cell of n's position in e's egocentric S/4 grid), gain
        max(coverage_n - coverage_e, 0) * 100.  Entries without a target are dropped.  Returns the number of records stored;
        ``store`` is a ReplayRing or any callable taking the record dict."""
        put = store.store_experience if hasattr(store, "store_experience") else store
        n_stored = 0
        E = self.entries
        for e in range(len(E)):
            pixels, gains = [], []
            for n in range(e + 1, len(E)):
                loc = self.transform(E[n][3][:3].unsqueeze(0), E[e][3], E[e][3].device)
                cell = self.cell_of(loc.squeeze(0), self.value_map_size, self.prediction_range)
                if 0 <= cell[0] < self.value_map_size[0] and 0 <= cell[1] < self.value_map_size[1]:
                    diff = E[n][0] - E[e][0]
                    gains.append(diff * 100 if diff > 0 else 0)
                    head = torch.as_tensor(E[n][4], device=cell.device).reshape(1).to(cell.dtype)
                    pixels.append(torch.cat((head, torch.stack((cell[0], cell[1])))))
            if pixels:
                put({"current_model_input": E[e][1], "current_gt_2d_layout": E[e][2],
                     "target_value_map_pixel": torch.stack(pixels, dim=0),
                     "actual_coverage_gain": torch.tensor(gains, dtype=torch.float32, device=pixels[0].device), "pose_i": pose_i})
                n_stored += 1
        self.entries = []
        return n_stored
