"""On-disk outputs of the analyze side: transition tables and transport maps round-trip through the
writers of spadot_b200.analyze in the format the reference's plot_OT reads (_analyze_utils.py:138,183) when anndata is
installed, and through the array fallback otherwise."""
import numpy as np
import pytest


def test_transition_table_round_trip_fallback_or_h5ad(tmp_path):
    from spadot_b200 import analyze
    tab = np.random.default_rng(0).uniform(0, 1, (4, 5))
    rows, cols = [f"0_{k}" for k in range(4)], [f"1_{k}" for k in range(5)]
    path = analyze.write_transition_table(str(tmp_path / "p_transition_table_0_1"), tab, rows, cols)
    assert path.endswith(".h5ad" if analyze.have_anndata() else ".npz")
    got, r, c = analyze.read_transition_table(path)
    assert r == rows and c == cols and np.array_equal(got, tab)
    prob = analyze.transition_probabilities(got)             # plot_OT's normalisation (_analyze_utils.py:185-193)
    assert prob.shape == tab.shape and np.all(prob <= 1.0 + 1e-12)


def test_h5ad_outputs_are_what_the_reference_reads(tmp_path):
    anndata = pytest.importorskip("anndata")
    from spadot_b200 import analyze
    tab = np.random.default_rng(1).uniform(0, 1, (3, 4))
    path = analyze.write_transition_table(str(tmp_path / "transition_table_0_1"), tab, ["0_0", "0_1", "0_2"], ["1_0", "1_1", "1_2", "1_3"])
    ad = anndata.read_h5ad(path)                             # _analyze_utils.py:183
    assert list(ad.obs_names) == ["0_0", "0_1", "0_2"] and list(ad.var_names) == ["1_0", "1_1", "1_2", "1_3"]
    assert np.allclose(ad.X, tab)
    plan = np.random.default_rng(2).uniform(0, 1, (6, 7))
    growth = np.ones((6, 4))
    p2 = analyze.write_transport_map(str(tmp_path / "OT_0_1"), plan, [f"c{i}" for i in range(6)], [f"d{j}" for j in range(7)], growth)
    ad2 = anndata.read_h5ad(p2)
    assert list(ad2.obs.columns) == ["g0", "g1", "g2", "g3"] and np.allclose(ad2.X, plan)
