"""CPU tests: the oracle against the golden vectors produced by the reference's own python
(oracle/make_golden.py), and the oracle's two routes against each other."""
import numpy as np
import pytest

from oracle import hiertcn_oracle as O
from helpers import GOLDEN, golden_variant_kwargs, load_hier_golden, small_case


ALL_GOLDEN = ["hier_default_arch", "hier_downsample_3lvl", "hier_gap_l2norm_warmstart"]


@pytest.mark.parametrize("name", ALL_GOLDEN)
def test_literal_oracle_matches_reference_graph(name):
    z, x, y, m, w = load_hier_golden(name)
    G = int(z["num_layer"])
    out = O.forward_loss_metrics(x, y, m, z["state0"], w, G, "f64", literal=True, **golden_variant_kwargs(z))
    np.testing.assert_allclose(out["pred"], z["pred_f64"], rtol=2e-6, atol=2e-6)   # fixture stored as f32
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out["loss"], z["loss_f64"], rtol=1e-12)
    np.testing.assert_allclose(out["loss_bt"], z["loss_bt_f64"], rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(out["ranks"], z["ranks_f64"])
    got = np.asarray([out[k] for k in ("recall1", "recall5", "recall10", "mrr", "mrp")])
    # loss.py:179 casts the rank comparison to tf.float32, so the reference's metric scalars carry
    # fp32 rounding even in the fp64 run of the fixture generator
    np.testing.assert_allclose(got, z["metrics_f64"], rtol=1e-6)


@pytest.mark.parametrize("name", ALL_GOLDEN)
def test_restructured_oracle_matches_reference_graph(name):
    """gather instead of one-hot matmul, hoisted GRU, split in-projection == the literal graph."""
    z, x, y, m, w = load_hier_golden(name)
    G = int(z["num_layer"])
    out = O.forward_loss_metrics(x, y, m, z["state0"], w, G, "f64", literal=False, **golden_variant_kwargs(z))
    np.testing.assert_allclose(out["pred"], z["pred_f64"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(out["loss"], z["loss_f64"], rtol=1e-11)
    np.testing.assert_array_equal(out["ranks"], z["ranks_f64"])


@pytest.mark.parametrize("name", ALL_GOLDEN)
def test_f32_restructured_within_1e4(name):
    """fp32 tier tolerance of the north star (1e-4 relative) holds for the fp32 oracle itself."""
    z, x, y, m, w = load_hier_golden(name)
    out = O.forward_loss_metrics(x, y, m, z["state0"], w, int(z["num_layer"]), "f32", **golden_variant_kwargs(z))
    assert abs(out["loss"] - z["loss_f64"]) <= 1e-4 * abs(z["loss_f64"])
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss"], z["loss_f32"], rtol=1e-5)


def test_bf16_tier_within_2e2():
    z, x, y, m, w = load_hier_golden("hier_default_arch")
    out = O.forward_loss_metrics(x, y, m, z["state0"], w, 2, "bf16")
    assert abs(out["loss"] - z["loss_f64"]) <= 2e-2 * abs(z["loss_f64"])


def test_embedding_gather_is_bit_exact_vs_onehot_matmul():
    rng = np.random.default_rng(0)
    E = rng.normal(size=(50, 128)).astype(np.float32)
    ids = rng.integers(0, 50, size=(4, 9))
    ids[0, :3] = 0
    lit = O.dense(O.one_hot_signed(ids, 50), E)
    np.testing.assert_array_equal(lit, O.emb_gather(ids, E))
    assert (O.emb_gather(ids, E)[ids == 0] == 0).all()


def test_loss_vectors():
    z = np.load(GOLDEN + "/loss_vectors.npz")
    for name in ("l2", "nce", "hinge_sigmoid", "hinge_logsigmoid", "hinge_linear", "bpr"):
        got = O.calc_loss_sampled(z["pred"], z["y"], z["y_impression"], name, int(z["num_neg_sample"]),
                                  float(z["nce_weight"]), float(z["hinge_delta"]))
        np.testing.assert_allclose(got, z["loss_" + name], rtol=1e-10, atol=1e-12, err_msg=name)
    for rm in ("l2", "inner_prod"):
        np.testing.assert_allclose(O.calc_score(z["pred"], z["y_impression"], rm), z["score_" + rm], rtol=1e-12)


def test_metric_fast_and_topk_vectors():
    z = np.load(GOLDEN + "/loss_vectors.npz")
    score, y_id = z["score"], z["y_id"]
    mask_y = np.sign(y_id).astype(np.float64)
    act = mask_y.sum(1)
    uc = np.sign(act).sum()
    act = act + 1e-6
    got = O.calc_metric_fast(score, mask_y, act, uc, y_id)
    np.testing.assert_allclose(np.asarray(got[:5]), z["metric_fast_scalars"], rtol=1e-6)
    np.testing.assert_array_equal(got[6], z["metric_fast_ranks"])
    np.testing.assert_allclose(got[5], z["metric_fast_ranks_float"], rtol=1e-6)
    v, i = O.top_k(score, score.shape[-1])
    np.testing.assert_array_equal(i, z["metric_topk_indices"])      # tie rule: lower index first
    np.testing.assert_array_equal(v, z["metric_topk_values"])


def test_causality_probe():
    """customized_tcn_cell.py:163-180: a spike at t=5 must not leak to t<5."""
    rng = np.random.default_rng(1)
    w = {}
    for lvl in range(3):
        w[f"t/temporal_conv_net/tblock_{lvl}/conv1/kernel"] = rng.normal(size=(4, 8, 8)).astype(np.float32)
        w[f"t/temporal_conv_net/tblock_{lvl}/conv1/bias"] = rng.normal(size=(8,)).astype(np.float32)
    x = rng.normal(size=(2, 30, 8)).astype(np.float32)
    x2 = x.copy()
    x2[:, 5, :] += 1000.0
    a = O.temporal_conv_net(x, w, "t", 3)
    b = O.temporal_conv_net(x2, w, "t", 3)
    np.testing.assert_array_equal(a[:, :5], b[:, :5])
    assert np.abs(a[:, 5:] - b[:, 5:]).max() > 0


def test_state_reset_and_row_independence():
    x, y, m, s0, w = small_case(B=4, S=3, L=5, N=31, seed=3)
    out = O.model_hier_restructured(x, y, m, s0, w, 2, "f32")
    # permuting users permutes outputs
    perm = np.array([2, 0, 3, 1])
    out_p = O.model_hier_restructured([a[perm] for a in x], [a[perm] for a in y], [a[perm] for a in m],
                                      s0[perm], w, 2, "f32")
    np.testing.assert_allclose(out[0][perm], out_p[0], rtol=1e-5, atol=1e-6)
    # mask 0 in the last slot zeroes the carried state
    m2 = [a.copy() for a in m]
    m2[-1][:] = 0
    assert (O.model_hier_restructured(x, y, m2, s0, w, 2, "f32")[1] == 0).all()


@pytest.mark.parametrize("literal", [True, False])
def test_torch_cpu_port_matches_golden(literal):
    """the CPU-baseline port that bench.py times is held to the same golden vectors"""
    from oracle.torch_cpu import CpuHierTCN
    z, x, y, m, w = load_hier_golden("hier_default_arch")
    out = CpuHierTCN(w).step(x, y, m, z["state0"], literal=literal, chunk=16)
    assert abs(out["loss"] - z["loss_f64"]) <= 1e-4 * abs(z["loss_f64"])
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss_bt"], z["loss_bt_f64"], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(out["ranks"], z["ranks_f64"])


def test_sign_bit_rank_is_strict_rank_up_to_one_ulp():
    """the rank count of the fused bf16 sweep (oracle restatement `rank_sign_bit` of k4_score_bf16.cu: kSignRank) against
    the strict compare of loss.py:179: never counts a logit <= the target, always counts a logit >= 2 ulp above it; only
    a logit exactly one ulp above the target may be missed"""
    rng = np.random.default_rng(0)
    zy = np.concatenate([rng.normal(size=4000) * 10.0 ** rng.integers(-3, 3, size=4000),
                         np.float32([1.0, -1.0, 0.75, 1.386, 1.387, -1.386, 2.0 ** -20, -(2.0 ** 20), 0.0])]).astype(np.float32)
    up = lambda a, k: a if k == 0 else up(np.nextafter(a, np.float32(np.inf)), k - 1)       # noqa: E731
    dn = lambda a, k: a if k == 0 else dn(np.nextafter(a, np.float32(-np.inf)), k - 1)      # noqa: E731
    # adversarial columns: the target itself, 1..3 ulp below, 1..3 ulp above
    cols = np.stack([zy, dn(zy, 1), dn(zy, 2), dn(zy, 3), up(zy, 1), up(zy, 2), up(zy, 3)], 1).astype(np.float32)
    for j in range(4):                                                 # <= target: never counted
        assert (O.rank_sign_bit(cols[:, j:j + 1], zy) == 0).all(), j
    for j in (5, 6):                                                   # >= 2 ulp above: always counted
        assert (O.rank_sign_bit(cols[:, j:j + 1], zy) == 1).all(), j
    one = O.rank_sign_bit(cols[:, 4:5], zy)                            # exactly 1 ulp above: either
    assert set(np.unique(one)) <= {0, 1} and 0 < one.mean() <= 1.0
    # random logits: strict - #[z == nextafter(z_y)] <= rank <= strict
    z = (rng.normal(size=(64, 5000)) * 3).astype(np.float32)
    y = rng.integers(0, 5000, size=64)
    zyr = z[np.arange(64), y]
    z[:, :8] = np.stack([up(zyr, k) for k in range(4)] + [dn(zyr, k) for k in range(4)], 1)   # ties and near-ties in every row
    strict = (z > zyr[:, None]).sum(1)
    one_ulp = (z == np.nextafter(zyr, np.float32(np.inf))[:, None]).sum(1)
    got = O.rank_sign_bit(z, zyr)
    assert ((got <= strict) & (got >= strict - one_ulp)).all()


def test_k4_exp2_polynomials_meet_their_stated_error():
    """the FMA-pipe exponentials of the bf16 catalog sweep (k4_score_bf16.cu: ex2_poly2): coefficients are read from the
    CUDA source, evaluated in float32 exactly like the kernel (round-to-nearest split, Horner, p(0) = 1) and compared
    with 2^t: degree 3 within 1.1e-4 relative (the default), degree 2 within 2.0e-3 (HTCN_K4_EPI=104)."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hiertcn_b200", "csrc",
                            "k4_score_bf16.cu")).read()
    body = src[src.index("__device__ __forceinline__ float2 ex2_poly2"):]
    body = body[:body.index("return make_float2")]
    deg2 = [np.float32(x) for x in re.findall(r"make_float2\((0\.\d+)f,", body[body.index("if (kDeg2)"):body.index("} else {", body.index("if (kDeg2)"))])]
    deg3 = [np.float32(x) for x in re.findall(r"make_float2\((0\.\d+)f,", body[body.index("} else {", body.index("if (kDeg2)")):])]
    assert len(deg2) == 2 and len(deg3) == 3, (deg2, deg3)
    t = np.linspace(-60.0, 60.0, 400001).astype(np.float32)
    magic = np.float32(12582912.0)
    r = (t + magic).astype(np.float32)
    n = (r - magic).astype(np.float32)
    f = (t - n).astype(np.float32)
    assert np.abs(f).max() <= 0.5

    def horner(cs):
        p = np.float32(cs[0]) * f + np.float32(cs[1])
        for c in cs[2:]:
            p = (p * f + np.float32(c)).astype(np.float32)
        return ((p * f).astype(np.float32) + np.float32(1.0)).astype(np.float32)

    exact = np.exp2(f.astype(np.float64))
    e3 = np.abs(horner(deg3).astype(np.float64) / exact - 1).max()
    e2 = np.abs(horner(deg2).astype(np.float64) / exact - 1).max()
    assert e3 <= 1.1e-4 and e2 <= 2.0e-3, (e3, e2)
    # 2^n by exponent-field addition is exact for |t| < 125 (the range the kernel's max|t| guard enforces)
    scaled = (horner(deg3).view(np.int32) + (r.view(np.int32) << 23)).view(np.float32)
    assert np.abs(scaled.astype(np.float64) / np.exp2(t.astype(np.float64)) - 1).max() <= 1.1e-4
