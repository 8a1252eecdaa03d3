"""Weight contract of the hot path: a state dict keyed by the TF variable names the reference graph
creates (derived from scope usage in model_hier.py:28,50,85, model_tcn.py:26,35,41,
customized_tcn_cell.py:80-87,153-154, customed_gru_cell.py:315,328,1056,1170-1185), so that
"identical weights" is well defined against a real TF-1.x run.  numpy only (host side)."""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np


def hier_weight_shapes(item_num: int, hidden_dim: int = 128, num_layer: int = 2,
                       tcn_channel=(128, 128), kernel_size: int = 5, emb_dim: int = 128,
                       output_dim: int | None = None) -> "OrderedDict[str, tuple]":
    out_dim = item_num if output_dim is None else output_dim
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["hier/emb/kernel"] = (item_num, emb_dim)
    s["hier/emb/bias"] = (emb_dim,)
    s["hier/tcn/emb/kernel"] = (emb_dim + num_layer * hidden_dim, 128)      # model_tcn.py:35 units=128
    cin = 128
    for lvl, c in enumerate(tcn_channel):
        p = f"hier/tcn/temporal_conv_net/tblock_{lvl}"
        s[p + "/conv1/kernel"] = (kernel_size, cin, c)
        s[p + "/conv1/bias"] = (c,)
        if cin != c:                                                        # customized_tcn_cell.py:102-106
            s[p + "/dense/kernel"] = (cin, c)
            s[p + "/dense/bias"] = (c,)
        cin = c
    s["hier/tcn/dense/kernel"] = (cin, out_dim)
    s["hier/tcn/dense/bias"] = (out_dim,)
    inp = emb_dim
    for g in range(num_layer):
        p = f"hier/multi_rnn_cell/cell_{g}/gru_cell"
        s[p + "/gates/kernel"] = (inp + hidden_dim, 2 * hidden_dim)
        s[p + "/gates/bias"] = (2 * hidden_dim,)
        s[p + "/candidate/kernel"] = (inp + hidden_dim, hidden_dim)
        s[p + "/candidate/bias"] = (hidden_dim,)
        inp = hidden_dim
    return s


def tcn_weight_shapes(in_dim: int, tcn_channel=(128, 128, 128, 128), kernel_size: int = 5,
                      output_dim: int | None = None, scope: str = "tcn") -> "OrderedDict[str, tuple]":
    """Single-level model_tcn (model_tcn.py:26-44), BASELINE config 3."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s[f"{scope}/emb/kernel"] = (in_dim, 128)
    cin = 128
    for lvl, c in enumerate(tcn_channel):
        p = f"{scope}/temporal_conv_net/tblock_{lvl}"
        s[p + "/conv1/kernel"] = (kernel_size, cin, c)
        s[p + "/conv1/bias"] = (c,)
        if cin != c:
            s[p + "/dense/kernel"] = (cin, c)
            s[p + "/dense/bias"] = (c,)
        cin = c
    if output_dim is not None:
        s[f"{scope}/dense/kernel"] = (cin, output_dim)
        s[f"{scope}/dense/bias"] = (output_dim,)
    return s


def _glorot_uniform(rng, shape):
    """[TF-sem] default initializer of tf.layers / get_variable: glorot_uniform."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:                                   # conv kernel [K, Cin, Cout]
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_weights(shapes, seed: int = 1234, kernel_scale: float = 1.0, bias_noise: float = 0.0):
    """glorot-uniform kernels, zero biases, GRU gate bias 1.0 (customed_gru_cell.py:312-321).

    ``kernel_scale`` > 1 gives the "trained-like" variant (logits not near-uniform);
    ``bias_noise`` > 0 perturbs biases so tests cannot pass with a dropped bias."""
    rng = np.random.default_rng(seed)
    w = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("/kernel"):
            w[name] = _glorot_uniform(rng, shape) * np.float32(kernel_scale)
        elif name.endswith("gates/bias"):
            w[name] = np.ones(shape, np.float32)
        else:
            w[name] = np.zeros(shape, np.float32)
        if bias_noise and name.endswith("/bias"):
            w[name] = w[name] + rng.normal(0, bias_noise, size=shape).astype(np.float32)
    return w


def weights_sha256(w) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k], dtype=np.float32).tobytes())
    return h.hexdigest()


def save_npz(path, w):
    np.savez(path, **{k.replace("/", "|"): v for k, v in w.items()})


def load_npz(path):
    z = np.load(path)
    return OrderedDict((k.replace("|", "/"), z[k]) for k in z.files)


def fold_weightnorm(w):
    """Weight-norm is a pure re-parameterisation (customized_dense_layer.py:140-142,
    customized_convolution_layer.py:146-148): fold ``g`` into the kernel at load time."""
    out = OrderedDict()
    for k, v in w.items():
        if k.endswith("/g"):
            continue
        if k.endswith("/kernel") and (k[:-len("kernel")] + "g") in w:
            g = w[k[:-len("kernel")] + "g"]
            if v.ndim == 3:
                nrm = np.sqrt(np.maximum((v * v).sum((0, 1), keepdims=True), 1e-12))
                v = g.reshape(1, 1, -1) * v / nrm
            else:
                v = g * v / np.sqrt((v * v).sum(0, keepdims=True))
        out[k] = v.astype(np.float32)
    return out


# --------------------------------------------------------------------------------------------------------------------
# Device layout.  The sm_100a kernels are built for 128-wide blocks (MMA tiles, 512-byte rows): every width of the model
# -- embedding (model_hier.py:50), GRU units (args.py:51), TCN channels (args.py:56) -- is run ZERO-PADDED to 128.
# A padded channel has zero weights and zero bias in front of a relu, so it stays exactly zero through every level; a
# padded GRU unit has zero weights and zero gate biases, so r = u = 1/2, c = 0 and h' = h/2 = 0 from a zero state; nothing
# reads a padded value through a non-zero weight.  Results equal the unpadded model's (the extra terms are exact zeros),
# and the gradients of the padded entries are exactly zero, so training never moves them.
# --------------------------------------------------------------------------------------------------------------------
PAD = 128


def _pad2(a, rows, cols):
    out = np.zeros((rows, cols), np.float32)
    out[:a.shape[0], :a.shape[1]] = a
    return out


def _pad1(a, n):
    out = np.zeros(n, np.float32)
    out[:a.shape[0]] = a
    return out


def layout_meta(w, scope="hier"):
    """shapes the TF-named dict ``w`` implies: emb_dim, hidden_dim, num_layer, channels, kernel_size, per-level
    down-sample flags.  ``scope`` 'hier' = model_hier (GRU + conditioned TCN), otherwise a plain model_tcn scope."""
    tcn = scope + "/tcn" if scope == "hier" else scope
    m = {"scope": scope, "tcn": tcn}
    chans, ds, l = [], [], 0
    while f"{tcn}/temporal_conv_net/tblock_{l}/conv1/kernel" in w:
        k = w[f"{tcn}/temporal_conv_net/tblock_{l}/conv1/kernel"]
        m["K"] = int(k.shape[0])
        chans.append(int(k.shape[2]))
        ds.append(f"{tcn}/temporal_conv_net/tblock_{l}/dense/kernel" in w)
        l += 1
    m["channels"], m["ds"] = chans, ds
    m.setdefault("K", 1)
    if scope == "hier":
        m["ed"] = int(w["hier/emb/kernel"].shape[1])
        G = 0
        while f"hier/multi_rnn_cell/cell_{G}/gru_cell/gates/kernel" in w:
            G += 1
        m["G"] = G
        m["H"] = int(w["hier/multi_rnn_cell/cell_0/gru_cell/candidate/kernel"].shape[1]) if G else PAD
    else:
        m["ed"], m["G"], m["H"] = PAD, 0, PAD
    for name, v in (("emb_dim", m["ed"]), ("hidden_dim", m["H"])):
        if v > PAD:
            raise NotImplementedError("%s = %d: widths above 128 are not built (the kernels run 128-wide blocks)" % (name, v))
    for c in chans:
        if c > 2 * PAD:
            raise NotImplementedError("tcn_channel = %d: at most 256 (two 128-wide planes, fp32 tier)" % c)
    # 128-wide planes per level output (1, or 2 for the 129..256-channel levels of args.py:310-311); the in-projection
    # always produces 128 channels (model_tcn.py:35)
    m["planes"] = [-(-c // PAD) for c in chans]
    m["wide"] = any(p > 1 for p in m["planes"])
    return m


def to_device_layout(w, scope="hier"):
    """TF-named weights -> dict of the arrays the kernels read, every width zero-padded to 128:
    E [N,128], b_emb [128], w_in_x [128,128], w_in_state [G*128,128], conv_w{l} [K,128,128], conv_b{l} [128],
    ds_w{l} [128,128] / ds_b{l} [128] (levels that change the width, customized_tcn_cell.py:102-106),
    gate_w{g} [256,256] (rows [x | h], columns [r | u]), gate_b{g} [256], cand_w{g} [256,128], cand_b{g} [128],
    w_out [128,N] (TF layout), b_out [N].  Returns (dict, meta)."""
    m = layout_meta(w, scope)
    tcn, ed, H, G = m["tcn"], m["ed"], m["H"], m["G"]
    d = OrderedDict()
    if scope == "hier":
        d["E"] = _pad2(w["hier/emb/kernel"], w["hier/emb/kernel"].shape[0], PAD)
        d["b_emb"] = _pad1(w["hier/emb/bias"], PAD)
        w_in = w[tcn + "/emb/kernel"]                                   # [ed + G*H, 128]
        d["w_in_x"] = _pad2(w_in[:ed], PAD, PAD)
        ws = np.zeros((max(G, 1) * PAD, PAD), np.float32)
        for g in range(G):
            ws[g * PAD:g * PAD + H] = w_in[ed + g * H:ed + (g + 1) * H]
        d["w_in_state"] = ws
    else:
        d["E"] = np.ascontiguousarray(w[tcn + "/emb/kernel"], np.float32)    # the 'emb' dense applied to a one-hot = a gather
        d["b_emb"] = np.zeros(PAD, np.float32)
        d["w_in_x"] = np.eye(PAD, dtype=np.float32)
    pin = 1
    for l, c in enumerate(m["channels"]):
        p = f"{tcn}/temporal_conv_net/tblock_{l}"
        k = w[p + "/conv1/kernel"]
        pout = m["planes"][l]
        if pin == 1 and pout == 1:
            kw = np.zeros((k.shape[0], PAD, PAD), np.float32)
            kw[:, :k.shape[1], :k.shape[2]] = k
        else:
            kw = _plane_blocks(k, pin, pout)           # [pout, pin*K, 128, 128]
        d[f"conv_w{l}"], d[f"conv_b{l}"] = kw, _pad1(w[p + "/conv1/bias"], pout * PAD)
        if m["ds"][l]:
            dk = w[p + "/dense/kernel"]
            d[f"ds_w{l}"] = _pad2(dk, PAD, PAD) if pin == 1 and pout == 1 else _plane_blocks(dk[None], pin, pout)
            d[f"ds_b{l}"] = _pad1(w[p + "/dense/bias"], pout * PAD)
        pin = pout
    for g in range(G):
        p = f"hier/multi_rnn_cell/cell_{g}/gru_cell"
        inp = ed if g == 0 else H
        gw, gb, cw, cb = w[p + "/gates/kernel"], w[p + "/gates/bias"], w[p + "/candidate/kernel"], w[p + "/candidate/bias"]
        GW, CW, GB = np.zeros((2 * PAD, 2 * PAD), np.float32), np.zeros((2 * PAD, PAD), np.float32), np.zeros(2 * PAD, np.float32)
        for r0, r1, dst in ((0, inp, 0), (inp, inp + H, PAD)):          # rows: input block, hidden block
            GW[dst:dst + r1 - r0, :H] = gw[r0:r1, :H]                   # r gate columns
            GW[dst:dst + r1 - r0, PAD:PAD + H] = gw[r0:r1, H:]          # u gate columns
            CW[dst:dst + r1 - r0, :H] = cw[r0:r1]
        GB[:H], GB[PAD:PAD + H] = gb[:H], gb[H:]
        d[f"gate_w{g}"], d[f"gate_b{g}"], d[f"cand_w{g}"], d[f"cand_b{g}"] = GW, GB, CW, _pad1(cb, PAD)
    if tcn + "/dense/kernel" in w:
        wo = w[tcn + "/dense/kernel"]
        d["w_out"] = _pad2(wo, pin * PAD, wo.shape[1])
        d["b_out"] = np.ascontiguousarray(w[tcn + "/dense/bias"], np.float32)
    return d, m


def _plane_blocks(k, pin, pout):
    """TF conv kernel [K, Cin, Cout] -> [pout, pin*K, 128, 128]: block (po, pi*K + tap) = k[tap, pi*128.., po*128..]
    zero-padded -- a level whose input has `pin` planes is run as pin*K taps over 128-wide planes"""
    K = k.shape[0]
    out = np.zeros((pout, pin * K, PAD, PAD), np.float32)
    for po in range(pout):
        for pi in range(pin):
            blk = k[:, pi * PAD:(pi + 1) * PAD, po * PAD:(po + 1) * PAD]
            out[po, pi * K:(pi + 1) * K, :blk.shape[1], :blk.shape[2]] = blk
    return out


def _unplane_blocks(b, K, cin, cout):
    pout, pin = b.shape[0], b.shape[1] // K
    k = np.zeros((K, pin * PAD, pout * PAD), np.float32)
    for po in range(pout):
        for pi in range(pin):
            k[:, pi * PAD:(pi + 1) * PAD, po * PAD:(po + 1) * PAD] = b[po, pi * K:(pi + 1) * K]
    return k[:, :cin, :cout]


def from_device_layout(d, m):
    """inverse of ``to_device_layout``: slices the padding off and restores the TF names / shapes.  ``d`` may hold the
    output table as ``w_out`` [128,N] or as ``wt`` [N,128] (its transpose, what the scoring kernels read)."""
    tcn, ed, H, G = m["tcn"], m["ed"], m["H"], m["G"]
    w = OrderedDict()
    if m["scope"] == "hier":
        w["hier/emb/kernel"] = d["E"][:, :ed]
        w["hier/emb/bias"] = d["b_emb"][:ed]
        w[tcn + "/emb/kernel"] = np.concatenate([d["w_in_x"][:ed]] + [d["w_in_state"][g * PAD:g * PAD + H] for g in range(G)], 0)
    else:
        w[tcn + "/emb/kernel"] = d["E"]
    cin = PAD
    for l, c in enumerate(m["channels"]):
        p = f"{tcn}/temporal_conv_net/tblock_{l}"
        cw = d[f"conv_w{l}"]
        w[p + "/conv1/kernel"] = cw[:, :cin, :c] if cw.ndim == 3 else _unplane_blocks(cw, m["K"], cin, c)
        w[p + "/conv1/bias"] = d[f"conv_b{l}"][:c]
        if m["ds"][l]:
            dw = d[f"ds_w{l}"]
            w[p + "/dense/kernel"] = dw[:cin, :c] if dw.ndim == 2 else _unplane_blocks(dw, 1, cin, c)[0]
            w[p + "/dense/bias"] = d[f"ds_b{l}"][:c]
        cin = c
    if "w_out" in d or "wt" in d:
        wo = d["w_out"] if "w_out" in d else np.ascontiguousarray(d["wt"].T)
        w[tcn + "/dense/kernel"], w[tcn + "/dense/bias"] = wo[:cin], d["b_out"]
    for g in range(G):
        p = f"hier/multi_rnn_cell/cell_{g}/gru_cell"
        inp = ed if g == 0 else H
        GW, GB, CW, CB = d[f"gate_w{g}"], d[f"gate_b{g}"], d[f"cand_w{g}"], d[f"cand_b{g}"]
        rows = np.r_[0:inp, PAD:PAD + H]
        w[p + "/gates/kernel"] = np.concatenate([GW[rows][:, :H], GW[rows][:, PAD:PAD + H]], 1)
        w[p + "/gates/bias"] = np.concatenate([GB[:H], GB[PAD:PAD + H]])
        w[p + "/candidate/kernel"], w[p + "/candidate/bias"] = CW[rows][:, :H], CB[:H]
    return OrderedDict((k, np.ascontiguousarray(v, dtype=v.dtype)) for k, v in w.items())


def pad_state(state, m):
    """carried user state [B, G*H] -> device layout [B, G*128]"""
    state = np.asarray(state, np.float32)
    H, G = m["H"], m["G"]
    if H == PAD:
        return state
    out = np.zeros((state.shape[0], G * PAD), np.float32)
    for g in range(G):
        out[:, g * PAD:g * PAD + H] = state[:, g * H:(g + 1) * H]
    return out


def unpad_state(state, m):
    H, G = m["H"], m["G"]
    if H == PAD:
        return state
    return np.concatenate([state[:, g * PAD:g * PAD + H] for g in range(G)], 1)
