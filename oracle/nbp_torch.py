"""oracle/nbp_torch.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch fp32 restatement of the reference network
``next_best_path/networks/nbp_model.py`` (NBP.forward :110-160, NBP.loss :162-173,
blocks :8-62), written functionally over a ``state_dict`` so that it can run on the
GPU box's host cores where /root/reference does not exist.

PINNED: ``tests/golden/make_golden.py`` ran the reference's own ``NBP`` class (imported
from /root/reference in the build container) and committed its outputs/gradients for a
seeded state_dict; ``tests/test_oracle_golden.py`` checks this file against those fixtures,
and -- when /root/reference is present -- against the live reference module bit-for-bit.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


_MOMENTUM = [BN_MOMENTUM]   # calibrate_bn() temporarily overrides this


def _bn(x, sd, prefix, training):
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        # F.batch_norm updates running stats in place exactly as nn.BatchNorm2d does
        out = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, _MOMENTUM[0], BN_EPS)
        sd[prefix + ".num_batches_tracked"] += 1
        return out
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], False, BN_MOMENTUM, BN_EPS)


def _conv(x, sd, prefix, pad):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], stride=1, padding=pad)


class Decisions:
    """The network's discrete decisions (which side of every ReLU an element falls on, which of its four inputs every
    2x2 max-pool forwards), keyed by the BatchNorm in front of the ReLU ("Conv1.conv.1", "Up5_1.up.2", ...), by
    "Att{l}_{d}.relu" for the attention gates and by "pool2".."pool5" (the pool feeding encoder level l).

    ``native`` is filled by ``forward`` with the decisions its own arithmetic takes.  If ``forced`` is given, the forward
    pass takes those decisions instead (ReLU: multiply by the given 0/1 mask, pool: gather the given window index, window
    order (0,0),(0,1),(1,0),(1,1)), which pins the evaluation to ONE linear piece of the piecewise-linear network: two
    arithmetics that take the same decisions differ by rounding only.  Test infrastructure for the gradient parity test
    (tests/test_train_gpu.py): the whole-network gradient is discontinuous across a decision flip, so a comparison is only
    a precision statement when both sides are on the same piece."""

    def __init__(self, forced=None):
        self.forced = forced
        self.native = {}


def _relu(x, name, dec):
    if dec is None:
        return F.relu(x)
    dec.native[name] = x.detach() > 0
    if dec.forced is None:
        return F.relu(x)
    return x * dec.forced[name].to(x.dtype)


def _windows(x):
    B, C, H, W = x.shape
    return x.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)


def first_argmax4(win):
    """Index of the first maximum over the last (window) dimension -- torch.max_pool2d's and the CUDA kernel's tie rule."""
    best = torch.zeros(win.shape[:-1], dtype=torch.long, device=win.device)
    bv = win[..., 0]
    for k in range(1, 4):
        upd = win[..., k] > bv
        best = torch.where(upd, torch.full_like(best, k), best)
        bv = torch.where(upd, win[..., k], bv)
    return best


def _pool(x, name, dec):
    if dec is None:
        return F.max_pool2d(x, 2, 2)
    win = _windows(x)
    dec.native[name] = first_argmax4(win.detach())
    idx = dec.native[name] if dec.forced is None else dec.forced[name]
    return win.gather(-1, idx.unsqueeze(-1)).squeeze(-1)


def _double_conv(x, sd, name, training, dec=None):
    """conv_block nbp_model.py:8-21: (3x3 conv + BN + ReLU) x 2 at Sequential indices 0,1 / 3,4."""
    x = _relu(_bn(_conv(x, sd, f"{name}.conv.0", 1), sd, f"{name}.conv.1", training), f"{name}.conv.1", dec)
    return _relu(_bn(_conv(x, sd, f"{name}.conv.3", 1), sd, f"{name}.conv.4", training), f"{name}.conv.4", dec)


def _up(x, sd, name, training, dec=None):
    """up_conv nbp_model.py:23-34: nearest x2, 3x3 conv, BN, ReLU (Sequential indices 1, 2)."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return _relu(_bn(_conv(x, sd, f"{name}.up.1", 1), sd, f"{name}.up.2", training), f"{name}.up.2", dec)


def _gate(g, x, sd, name, training, dec=None):
    """Attention_block nbp_model.py:36-62: x * sigmoid(BN(psi(relu(BN(Wg g) + BN(Wx x)))))."""
    g1 = _bn(_conv(g, sd, f"{name}.W_g.0", 0), sd, f"{name}.W_g.1", training)
    x1 = _bn(_conv(x, sd, f"{name}.W_x.0", 0), sd, f"{name}.W_x.1", training)
    a = _relu(g1 + x1, f"{name}.relu", dec)
    psi = torch.sigmoid(_bn(_conv(a, sd, f"{name}.psi.0", 0), sd, f"{name}.psi.1", training))
    return x * psi


def forward(sd, x, training: bool = False, decisions: "Decisions | None" = None):
    """NBP.forward nbp_model.py:110-160. ``sd`` maps the reference's state_dict keys to tensors
    (parameters may require grad).  Returns (out1 (B,8,S/4,S/4), out2 (B,1,S,S)).  ``decisions``: see ``Decisions``
    (None = the reference's own F.relu / F.max_pool2d calls)."""
    dec = decisions
    _dc = _double_conv
    _double_conv_d = lambda t, s, n, tr: _dc(t, s, n, tr, dec)
    x1 = _double_conv_d(x, sd, "Conv1", training)
    x2 = _double_conv_d(_pool(x1, "pool2", dec), sd, "Conv2", training)
    x3 = _double_conv_d(_pool(x2, "pool3", dec), sd, "Conv3", training)
    x4 = _double_conv_d(_pool(x3, "pool4", dec), sd, "Conv4", training)
    x5 = _double_conv_d(_pool(x4, "pool5", dec), sd, "Conv5", training)

    def stage(d, skip, lvl, dec_no):
        u = _up(d, sd, f"Up{lvl}_{dec_no}", training, dec)
        s = _gate(u, skip, sd, f"Att{lvl}_{dec_no}", training, dec)
        return _double_conv_d(torch.cat((s, u), dim=1), sd, f"Up_conv{lvl}_{dec_no}", training)

    d = stage(x5, x4, 5, 1)
    d = stage(d, x3, 4, 1)
    out1 = _conv(d, sd, "Final1", 0)

    d = stage(x5, x4, 5, 2)
    d = stage(d, x3, 4, 2)
    d = stage(d, x2, 3, 2)
    d = stage(d, x1, 2, 2)
    out2 = torch.sigmoid(_conv(d, sd, "Final2.0", 0))
    return out1, out2


def loss(sd, pred1, target1, pred2, target2):
    """NBP.loss nbp_model.py:162-173 (homoscedastic-uncertainty weighting of MSE + BCE)."""
    lv = sd["log_vars"]
    s1, s2 = torch.exp(2 * lv[0]), torch.exp(2 * lv[1])
    l1 = (1.0 / (2.0 * s1)) * F.mse_loss(pred1, target1) + lv[0]
    l2 = (1.0 / s2) * F.binary_cross_entropy(pred2, target2) + lv[1]
    return l1 + l2


# ----------------------------------------------------------------------------- weights
_ENC = [("Conv1", 5, 64), ("Conv2", 64, 128), ("Conv3", 128, 256), ("Conv4", 256, 512), ("Conv5", 512, 1024)]
_DEC = {1: [5, 4], 2: [5, 4, 3, 2]}
_CH = {5: 1024, 4: 512, 3: 256, 2: 128, 1: 64}


def state_dict_spec(img_ch: int = 5, out1: int = 8, out2: int = 1):
    """Ordered (key, shape, kind) list of the reference's 327 state_dict entries
    (``log_vars`` first: root parameters precede child modules)."""
    spec = [("log_vars", (2,), "param")]

    def conv(prefix, cin, cout, k):
        spec.append((prefix + ".weight", (cout, cin, k, k), "conv_w"))
        spec.append((prefix + ".bias", (cout,), "conv_b"))

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
                     (prefix + ".running_mean", (c,), "bn_rm"), (prefix + ".running_var", (c,), "bn_rv"),
                     (prefix + ".num_batches_tracked", (), "bn_n")])

    def block(name, cin, cout):
        conv(f"{name}.conv.0", cin, cout, 3); bn(f"{name}.conv.1", cout)
        conv(f"{name}.conv.3", cout, cout, 3); bn(f"{name}.conv.4", cout)

    for name, cin, cout in _ENC:
        block(name, img_ch if name == "Conv1" else cin, cout)
    for dec in (1, 2):
        for lvl in _DEC[dec]:
            cin, cout = _CH[lvl], _CH[lvl] // 2
            conv(f"Up{lvl}_{dec}.up.1", cin, cout, 3); bn(f"Up{lvl}_{dec}.up.2", cout)
            fint = cout // 2
            conv(f"Att{lvl}_{dec}.W_g.0", cout, fint, 1); bn(f"Att{lvl}_{dec}.W_g.1", fint)
            conv(f"Att{lvl}_{dec}.W_x.0", cout, fint, 1); bn(f"Att{lvl}_{dec}.W_x.1", fint)
            conv(f"Att{lvl}_{dec}.psi.0", fint, 1, 1); bn(f"Att{lvl}_{dec}.psi.1", 1)
            block(f"Up_conv{lvl}_{dec}", cin, cout)
        if dec == 1:
            conv("Final1", 256, out1, 1)
    conv("Final2.0", 64, out2, 1)
    return spec


def seeded_state_dict(seed: int = 9, head_scale: float = 1.0):
    """A deterministic, well-conditioned weight set that does NOT depend on torch's module
    initialisers (so fixtures stay valid across torch versions): conv weights ~ U(-b, b) with
    b = sqrt(3 / fan_in) (unit-gain), small biases, BN gamma in [0.8, 1.2], beta in [-0.1, 0.1],
    running_mean in [-0.2, 0.2], running_var in [0.6, 1.4].  ``head_scale`` multiplies Final1 so that
    out1 is O(1-10) like the x100 coverage-gain targets (nbp_utils.py:668)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    u = lambda shape, lo, hi: torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo
    for key, shape, kind in state_dict_spec():
        if kind == "param":
            sd[key] = torch.zeros(shape)
        elif kind == "conv_w":
            fan_in = shape[1] * shape[2] * shape[3]
            b = (3.0 / fan_in) ** 0.5
            sd[key] = u(shape, -b, b)
        elif kind == "conv_b":
            sd[key] = u(shape, -0.05, 0.05)
        elif kind == "bn_w":
            sd[key] = u(shape, 0.8, 1.2)
        elif kind == "bn_b":
            sd[key] = u(shape, -0.1, 0.1)
        elif kind == "bn_rm":
            sd[key] = u(shape, -0.2, 0.2)
        elif kind == "bn_rv":
            sd[key] = u(shape, 0.6, 1.4)
        elif kind == "bn_n":
            sd[key] = torch.tensor(0, dtype=torch.long)
    sd["Final1.weight"] = sd["Final1.weight"] * head_scale
    return sd


def count_like_input(B: int, S: int, seed: int = 8, density: float = 0.04, rate: float = 25.0, n_traj: int = 20):
    """Synthetic model input shaped like nbp_planning.py:126-132: 4 sparse count images (walls seen
    top-down are thin and dense) + a sparse 0/1 trajectory image.  Integer-valued fp32."""
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(B, 5, S, S)
    occ = (torch.rand(B, 4, S, S, generator=g) < density).float()
    x[:, :4] = occ * (1.0 + torch.poisson(torch.full((B, 4, S, S), rate), generator=g))
    idx = torch.randint(0, S, (B, n_traj, 2), generator=g)
    for b in range(B):
        x[b, 4, idx[b, :, 0], idx[b, :, 1]] = 1.0
    return x


def calibrate_bn(sd, x):
    """Set every BatchNorm's running statistics to the batch statistics of ``x`` (one train-mode
    pass with momentum 1), so that eval-mode activations are O(1) (default
    running stats make value-map parity trivially small)."""
    _MOMENTUM[0] = 1.0
    try:
        with torch.no_grad():
            forward(sd, x, training=True)
    finally:
        _MOMENTUM[0] = BN_MOMENTUM
    for k in sd:
        if k.endswith("num_batches_tracked"):
            sd[k].zero_()
    return sd


def golden_state_dict(seed: int = 9, calib_S: int = 64, calib_B: int = 2):
    """The weight set every parity test and the bench use: seeded, BN-calibrated on count-like
    inputs, Final1 scaled so that out1 is O(1-10)."""
    sd = seeded_state_dict(seed)
    calibrate_bn(sd, count_like_input(calib_B, calib_S, seed=seed + 1))
    with torch.no_grad():
        o1, _ = forward(sd, count_like_input(1, calib_S, seed=seed + 2))
        scale = 5.0 / float(o1.abs().max().clamp_min(1e-6))
    sd["Final1.weight"] = sd["Final1.weight"] * scale
    sd["Final1.bias"] = sd["Final1.bias"] * scale
    return sd
