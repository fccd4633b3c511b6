"""CPU: oracle (oracle/schedule_oracle.py) and product host logic vs the reference's golden vectors / KATs.
Integer logic (space_timesteps, timestep_map, _WrappedModel gather) and fp64 tables must be bit-exact."""
import ast

import numpy as np
import pytest

import b200mdm
from b200mdm.diffusion import gaussian_diffusion as gd
from b200mdm.diffusion import respace as rs
from oracle import schedule_oracle as so

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
          "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
          "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
          "posterior_mean_coef1", "posterior_mean_coef2"]


def test_kat_constants_from_survey():
    # SURVEY.md section 8a rows a1/a2 (probed from the reference)
    b50 = so.named_betas("cosine", 50)
    assert b50[0] == 0.0017475135338653747 and b50[1] == 0.003688954726538851
    assert b50[48] == 0.7497570035157484 and b50[49] == 0.999
    assert so.named_betas("cosine", 1000)[0] == 4.128422482196914e-05
    assert so.named_betas("cosine", 10)[0] == 0.02790726288603096
    t = so.diffusion_tables(b50)
    assert t["alphas_cumprod"][0] == 0.9982524864661346 and t["alphas_cumprod"][49] == 9.71193029871257e-07
    assert t["posterior_mean_coef1"][0] == 1.0 and t["posterior_mean_coef2"][0] == 0.0
    assert t["posterior_mean_coef1"][49] == 0.03113283632546369 and t["posterior_mean_coef2"][49] == 0.031592095463485986
    assert t["posterior_log_variance_clipped"][0] == t["posterior_log_variance_clipped"][1] == -6.7361613348922935
    assert t["posterior_log_variance_clipped"][49] == -0.0019711940834747413
    # product side computes the same bits
    assert np.array_equal(gd.get_named_beta_schedule("cosine", 50), b50)


def test_space_timesteps_kats():
    for fn in (so.space_timesteps, rs.space_timesteps):
        s = sorted(fn(300, [10, 15, 20]))
        assert len(s) == 45 and s[:3] == [0, 11, 22] and 99 in s and 100 in s and 107 in s
        assert fn(1000, "ddim50") == set(range(0, 1000, 20))
        s = sorted(fn(1000, "50"))
        assert s[:6] == [0, 20, 41, 61, 82, 102] and s[-3:] == [958, 979, 999]
        with pytest.raises(ValueError):
            fn(1000, "ddim333")
        with pytest.raises(ValueError):
            fn(10, [11])
        assert fn(50, [50]) == set(range(50))


def _cases(g):
    i = 0
    while "case%d_meta" % i in g:
        sched, steps, resp = [str(v) for v in g["case%d_meta" % i]]
        yield i, sched, int(steps), ast.literal_eval(resp)
        i += 1


def test_tables_bit_exact_vs_reference_golden(golden):
    g = golden("schedule.npz")
    n = 0
    for i, sched, steps, resp in _cases(g):
        base = g["case%d_base_betas" % i]
        # oracle
        ob = so.named_betas(sched, steps)
        assert np.array_equal(ob, base)
        nb, tmap, orig = so.respaced(ob, so.space_timesteps(steps, resp))
        assert np.array_equal(np.array(tmap), g["case%d_timestep_map" % i]) and orig == steps
        ot = so.diffusion_tables(nb)
        # product
        d = rs.SpacedDiffusion(use_timesteps=rs.space_timesteps(steps, resp), betas=gd.get_named_beta_schedule(sched, steps),
                               model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                               loss_type=gd.LossType.MSE, rescale_timesteps=False)
        assert d.timestep_map == list(g["case%d_timestep_map" % i]) and d.original_num_steps == steps
        assert d.num_timesteps == len(tmap)
        for name in TABLES:
            ref = g["case%d_%s" % (i, name)]
            assert np.array_equal(ot[name], ref), (i, name, "oracle")
            assert np.array_equal(getattr(d, name), ref), (i, name, "product")
        n += 1
    assert n >= 7


def test_space_timesteps_vs_reference_golden(golden):
    g = golden("schedule.npz")
    i = 0
    while "space%d_args" % i in g:
        n, sc = [str(v) for v in g["space%d_args" % i]]
        sc = ast.literal_eval(sc)
        want = list(g["space%d_steps" % i])
        assert sorted(so.space_timesteps(int(n), sc)) == want
        assert sorted(rs.space_timesteps(int(n), sc)) == want
        i += 1
    assert i >= 10
    for e in g["space_errors"]:
        n, sc = str(e).split("|")
        with pytest.raises(ValueError):
            rs.space_timesteps(int(n), ast.literal_eval(sc))
        with pytest.raises(ValueError):
            so.space_timesteps(int(n), ast.literal_eval(sc))


def test_wrapped_model_gather(golden):
    import torch
    g = golden("schedule.npz")
    tm = sorted(rs.space_timesteps(1000, "50"))
    assert np.array_equal(so.wrapped_timesteps(tm, g["wrapped_in"]), g["wrapped_out"])
    wm = rs._WrappedModel(lambda x, ts, **kw: ts, tm, False, 1000)
    assert np.array_equal(wm(None, torch.from_numpy(g["wrapped_in"])).numpy(), g["wrapped_out"])


def test_schedule_rows_match_fp32_cast_of_tables():
    d = b200mdm.create_gaussian_diffusion(__import__("conftest").default_args())
    rows = d.schedule_rows(eta=0.0)
    assert rows.dtype == np.float32 and rows.shape == (50, 8)
    assert np.array_equal(rows[:, 0], d.posterior_mean_coef1.astype(np.float32))
    assert np.array_equal(rows[:, 1], d.posterior_mean_coef2.astype(np.float32))
    assert rows[0, 2] == 0.0 and rows[0, 7] == 0.0                     # no noise at t == 0
    want = np.exp(np.float32(0.5) * d.posterior_log_variance_clipped.astype(np.float32))
    assert np.array_equal(rows[1:, 2], want[1:])
    assert np.all(rows[:, 7] == 0.0)                                   # eta = 0
    rows2 = d.schedule_rows(eta=0.5)
    assert np.all(rows2[1:, 7] > 0) and np.all(np.isfinite(rows2))


def test_positional_table_matches_oracle():
    from b200mdm.model.mdm import positional_table
    import torch
    assert torch.equal(positional_table(300, 512), so.positional_table(300, 512))
