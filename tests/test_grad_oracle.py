"""CPU tests pinning the gradient / optimiser oracle (oracle/grad_oracle.py): its forward value against the golden
vectors made by the reference's own python, literal vs restructured gradients, finite differences, Adam."""
import numpy as np
import pytest

from oracle import grad_oracle as GO
from helpers import load_hier_golden, small_case


@pytest.mark.parametrize("literal", [True, False])
def test_loss_value_matches_reference_golden(literal):
    z, x, y, m, w = load_hier_golden("hier_default_arch")
    loss, g, state = GO.loss_and_grads(w, x, y, m, z["state0"], int(z["num_layer"]), literal=literal)
    np.testing.assert_allclose(loss, z["loss_f64"], rtol=1e-11)
    np.testing.assert_allclose(state, z["state_f64"], rtol=1e-10, atol=1e-12)
    assert set(g) == set(w)


def test_literal_and_restructured_gradients_agree():
    x, y, m, s0, w = small_case(B=4, S=3, L=6, N=61, seed=3)
    _, ga, _ = GO.loss_and_grads(w, x, y, m, s0, literal=True)
    _, gb, _ = GO.loss_and_grads(w, x, y, m, s0, literal=False)
    for k in w:
        np.testing.assert_allclose(ga[k], gb[k], rtol=1e-9, atol=1e-13, err_msg=k)
    assert np.all(ga["hier/emb/kernel"][0] == 0)            # id 0 never reads row 0 (model.py:59-61)


def test_gradients_match_finite_differences():
    import torch
    x, y, m, s0, w = small_case(B=3, S=2, L=5, N=37, seed=5)
    _, g, _ = GO.loss_and_grads(w, x, y, m, s0)
    rng = np.random.default_rng(0)

    def f(wd):
        p = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in wd.items()}
        return float(GO.loss_fp64(p, x, y, m, s0)[0])

    used = np.unique(np.concatenate([np.asarray(v).ravel() for v in y]))
    used = used[used > 0]
    for k in w:
        a = np.asarray(w[k], np.float64)
        for _ in range(3):
            idx = tuple(int(rng.integers(0, n)) for n in a.shape)
            if k == "hier/emb/kernel":
                idx = (int(rng.choice(used)),) + idx[1:]
            wp = {kk: np.asarray(vv, np.float64).copy() for kk, vv in w.items()}
            wm = {kk: np.asarray(vv, np.float64).copy() for kk, vv in w.items()}
            h = 1e-5
            wp[k][idx] += h
            wm[k][idx] -= h
            fd = (f(wp) - f(wm)) / (2 * h)
            assert abs(fd - g[k][idx]) <= 1e-6 * max(1.0, abs(fd)) + 1e-8, (k, idx, fd, g[k][idx])


def test_adam_tf_first_step_and_epsilon_placement():
    w = {"a": np.array([1.0, -2.0, 3.0])}
    g = {"a": np.array([0.5, -0.25, 0.0])}
    m = {"a": np.zeros(3)}
    v = {"a": np.zeros(3)}
    GO.adam_tf(w, g, m, v, 1, lr=0.01)
    # first step: m_hat/sqrt(v_hat) = sign(g), so |delta| ~= lr wherever g != 0; zero gradient -> no move
    np.testing.assert_allclose(w["a"], [1.0 - 0.01, -2.0 + 0.01, 3.0], rtol=0, atol=1e-6)
    # epsilon sits outside the bias-corrected root: with a tiny gradient the step shrinks by sqrt(1-b2) * |g| / eps
    w2, g2 = {"a": np.array([0.0])}, {"a": np.array([1e-9])}
    m2, v2 = {"a": np.zeros(1)}, {"a": np.zeros(1)}
    GO.adam_tf(w2, g2, m2, v2, 1, lr=0.01)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = -lr_t * (0.1 * 1e-9) / (np.sqrt(0.001 * 1e-18) + 1e-8)
    np.testing.assert_allclose(w2["a"][0], want, rtol=1e-12)


def test_training_reduces_loss():
    x, y, m, s0, w = small_case(B=4, S=2, L=5, N=41, seed=7, kernel_scale=1.0)
    batches = [(x, y, m)] * 4
    losses, w2, _ = GO.train_steps(w, batches, s0, lr=1e-2)
    assert losses[-1] < losses[0]


def test_sampled_loss_rows_of_the_autograd_restatement_match_the_numpy_oracle():
    """oracle/grad_oracle._sampled_rows (torch, differentiable) == oracle.calc_loss_sampled (numpy), which is pinned against the
    reference's own loss.py through tests/golden/loss_vectors.npz"""
    import torch
    from oracle import grad_oracle as GO
    from oracle import hiertcn_oracle as O
    rng = np.random.default_rng(3)
    B, T, k, N, d = 3, 4, 6, 40, 16
    table = rng.normal(size=(N, d))
    hout = rng.normal(size=(B, T, d))
    y_id = rng.integers(1, N, size=(B, T))
    neg = rng.integers(0, N, size=(B, T, k))
    tz = table.copy()
    tz[0] = 0.0                                                # id 0 -> zero row
    for kind in ("nce", "hinge_sigmoid", "hinge_logsigmoid", "hinge_linear", "bpr"):
        got = GO._sampled_rows(torch.tensor(hout), torch.tensor(table), torch.tensor(y_id), torch.tensor(neg), kind,
                               0.1, 1.0, 20).numpy()
        ref = O.calc_loss_sampled(hout, tz[y_id], tz[neg], kind, 20, 1.0, 0.1)
        np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
