#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference checkout.

Run ONCE in the development container (needs /root/reference, which does not
exist on the GPU box):  python tests/golden/make_golden.py

What it does
------------
1. Stubs the third-party modules the reference imports but that are absent here
   (ttvfast, h5py, ptemcee, seaborn, corner, matplotlib, mpl_toolkits) and imports the
   UNMODIFIED reference package from /root/reference.
2. Rebuilds the PlanetarySystem objects of the reference's own notebooks with the
   reference's own classes (SetPlanet / PlanetarySystem) and freezes every attribute
   the hot path consumes (planetarysystem.py:184-350) to JSON:
     doc2   Documentation/2- Setup.ipynb cells 26-47   (t0=0, ftime=221, dt=0.18)
     sys3   Examples/simple_fit.ipynb cells 6-12       (t0=0, ftime=500)
     sys3f  same planets, default simulation()          (full 1359 d tables)
3. Extracts the known answers stored in the notebooks' outputs (produced by the real
   ttvfast 0.3.0 on the author's machine): G1/G2 ephemerides, G1/G2/G3/G4 chi2.
4. Runs the reference's pure-Python plumbing (cube_to_physical, _insert_constants,
   intervals) on seeded random inputs and stores input/output pairs.
"""
import ast
import json
import os
import sys
import types

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    names = ["ttvfast", "ttvfast.models", "h5py", "ptemcee", "seaborn", "corner",
             "matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.gridspec",
             "mpl_toolkits", "mpl_toolkits.axes_grid1", "sklearn", "sklearn.linear_model"]
    for n in names:
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = lambda *a, **k: None
    sys.modules["sklearn.linear_model"].LinearRegression = object
    sys.modules["matplotlib.ticker"].FormatStrFormatter = object


def _nb_cells(path):
    with open(path) as fh:
        return json.load(fh)["cells"]


def _out_text(cell):
    chunks = []
    for o in cell.get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain") or ""
        chunks.append("".join(t))
    return "\n".join(chunks)


def _freeze_system(PS):
    import numpy as np
    ids = list(PS.planets_IDs.keys())
    obs = {}
    for pid in PS.transit_times:
        obs[pid] = {
            "planet_index": PS.planets_IDs[pid],
            "epochs": [int(e) for e in PS.transit_times[pid].keys()],
            "times": [float(t) for t in PS.transit_times[pid].values()],
            "sigma": [float(s) for s in PS.sigma_obs[pid].values()],
        }
    return {
        "system_name": PS.system_name,
        "mstar": float(PS.mstar), "rstar": float(PS.rstar), "rstarAU": float(PS.rstarAU),
        "t0": float(PS.t0), "ftime": float(PS.ftime), "dt": float(PS.dt),
        "npla": int(PS.npla), "ndim": int(PS.ndim),
        "planets_IDs": {k: int(v) for k, v in PS.planets_IDs.items()},
        "bounds": [[float(a), float(b)] for a, b in PS.bounds],
        "bi": [float(v) for v in PS.bi], "bf": [float(v) for v in PS.bf],
        "hypercube": [[float(a), float(b)] for a, b in PS.hypercube],
        "constant_params": {str(int(k)): float(v) for k, v in PS.constant_params.items()},
        "params_names": PS.params_names, "params_names_all": PS.params_names_all,
        "observed": obs,
        "second_term_logL": {k: float(v) for k, v in PS.second_term_logL.items()},
        "second_term_sum": float(sum(PS.second_term_logL.values())),
        "planet_order": ids,
    }


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    import numpy as np
    import nauyaca as nau
    from nauyaca import utils as rutils

    # ---- doc2: Documentation/2- Setup.ipynb cells 26-47
    d = os.path.join(REF, "Documentation", "inputs")
    Pb = nau.SetPlanet("Planet-b")
    Pb.load_ttvs(os.path.join(d, "planetb_ttvs.dat"))
    Pb.mass = [1, 50.]; Pb.period = [5.35, 5.36]; Pb.ecc = [0., 0.2]
    Pb.inclination = [90, 90]; Pb.ascending_node = [88.36, 88.36]
    Pc = nau.SetPlanet("Planet-c")
    Pc.load_ttvs(os.path.join(d, "planetc_ttvs.dat"))
    Pc.mass = [1, 50]; Pc.period = [11.83, 11.84]; Pc.ecc = [0, .2]
    Pc.inclination = [90, 90]; Pc.ascending_node = [70, 110.]
    doc2 = nau.PlanetarySystem("SystemX", mstar=0.9132, rstar=0.8632)
    doc2.add_planets([Pb, Pc])
    doc2.simulation(t0=0.0, ftime=221, dt=0.18)

    # ---- sys3: Examples/simple_fit.ipynb cells 6-12
    def three(ftime):
        e = os.path.join(REF, "Examples", "inputs")
        P1 = nau.SetPlanet('Planet-b')
        P1.mass = [0.001, 10]; P1.period = [10.47, 10.48]; P1.ecc = [0.0, 0.1]
        P1.inclination = [90, 90]; P1.ascending_node = [60, 120]
        P1.load_ttvs(os.path.join(e, "3pl_planet1_ttvs.dat"))
        P2 = nau.SetPlanet('Planet-c')
        P2.mass = [0.001, 10]; P2.period = [14.10, 14.11]; P2.ecc = [0.0, 0.1]
        P2.inclination = [85, 95]; P2.ascending_node = [60, 120]
        P2.load_ttvs(os.path.join(e, "3pl_planet2_ttvs.dat"))
        P3 = nau.SetPlanet('Planet-d')
        P3.mass = [0.001, 10]; P3.period = [23.57, 23.58]; P3.ecc = [0.0, 0.1]
        P3.inclination = [90.3, 90.3]; P3.ascending_node = [88.5, 88.5]
        P3.load_ttvs(os.path.join(e, "3pl_planet3_ttvs.dat"))
        PS = nau.PlanetarySystem("MySystem", mstar=0.522, rstar=0.4422)
        PS.add_planets([P1, P2, P3])
        if ftime is not None:
            PS.simulation(t0=0, ftime=ftime)
        else:
            PS.simulation(t0=0)
        return PS
    sys3 = three(500)
    sys3f = three(None)

    systems = {"doc2": _freeze_system(doc2), "sys3": _freeze_system(sys3),
               "sys3f": _freeze_system(sys3f)}
    with open(os.path.join(HERE, "systems.json"), "w") as fh:
        json.dump(systems, fh, indent=1)

    # ---- known answers from Examples/miscellany.ipynb
    cells = _nb_cells(os.path.join(REF, "Examples", "miscellany.ipynb"))
    # G1 inputs: cell 10 prints the 18 free physical parameters at full repr precision
    phys = [float(line.split(":")[1]) for line in _out_text(cells[10]).strip().splitlines()]
    assert len(phys) == 18
    g1_flat = [float(v) for v in rutils._insert_constants(sys3, phys)]
    eph1 = ast.literal_eval(_out_text(cells[33]).split("Transit ephemeris:")[1].strip())
    eph1b = ast.literal_eval(_out_text(cells[35]).split("Transit ephemeris:")[1].strip())
    assert eph1 == eph1b
    chi2_old_g1 = float(_out_text(cells[19]).strip())
    chi2_ind_old_g1 = ast.literal_eval(_out_text(cells[23]).strip())
    logl_old_g1 = float(_out_text(cells[26]).strip())
    # G2: cell 43 proposal (inc2 = 95 deg) -> cell 47 ephemeris, cell 49 chi2
    src43 = "".join(cells[43]["source"])
    prop = ast.literal_eval(src43.split("my_proposal =")[1].strip())
    g2_flat = [float(v) for v in rutils._insert_constants(sys3, prop)]
    eph2 = ast.literal_eval(_out_text(cells[47]).strip())
    chi2_g2 = float(_out_text(cells[49]).strip())
    # G3: cell 53 proposal (inc2 = 90 deg) -> chi2 only
    src53 = "".join(cells[53]["source"])
    prop3 = ast.literal_eval(src53.split("my_proposal =")[1].split("]")[0] + "]")
    g3_flat = [float(v) for v in rutils._insert_constants(sys3, prop3)]
    chi2_old_g3 = float(_out_text(cells[53]).strip())

    # G4: Documentation/3- Running Optimizers.ipynb (best optimizer solution, 9 digits)
    cells3 = _nb_cells(os.path.join(REF, "Documentation", "3- Running Optimizers.ipynb"))
    g4 = None
    for c in cells3:
        t = _out_text(c)
        if "Best chi2:" in t and "Best solution:" in t:
            chi2_old_g4 = float(t.split("Best chi2:")[1].split("Best solution:")[0].strip())
            arr = t.split("Best solution:")[1].replace("[", " ").replace("]", " ").split()
            g4 = [float(v) for v in arr]
            break
    assert g4 is not None and len(g4) == 11
    g4_flat = [float(v) for v in rutils._insert_constants(doc2, g4)]

    def eph_json(e):
        return {pid: {"epochs": [int(k) for k in d.keys()], "times": [float(v) for v in d.values()]}
                for pid, d in e.items()}

    # sigma convention: the notebooks predate planetarysystem.py:325 (sqrt(lo^2+hi^2)); all
    # shipped tables have lo == hi so current chi2 == notebook chi2 / 2 exactly.
    known = {
        "G1": {"system": "sys3", "flat_params": g1_flat, "phys_free": phys,
               "ephemeris": eph_json(eph1),
               "chi2_notebook_old_sigma": chi2_old_g1, "chi2_current": chi2_old_g1 / 2.0,
               "chi2_individual_old_sigma": chi2_ind_old_g1,
               "logl_notebook_old_sigma": logl_old_g1},
        "G2": {"system": "sys3", "flat_params": g2_flat, "ephemeris": eph_json(eph2),
               "chi2_current": chi2_g2},
        "G3": {"system": "sys3", "flat_params": g3_flat,
               "chi2_notebook_old_sigma": chi2_old_g3, "chi2_current": chi2_old_g3 / 2.0},
        "G4": {"system": "doc2", "flat_params": g4_flat, "phys_free": g4,
               "chi2_notebook_old_sigma": chi2_old_g4, "chi2_current": chi2_old_g4 / 2.0,
               "note": "inputs rounded to 9 significant digits in the notebook"},
    }
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(known, fh, indent=1)

    # ---- plumbing fixtures: reference cube_to_physical / intervals on seeded inputs
    rng = np.random.default_rng(20240607)
    plumb = {}
    for name, PS in (("doc2", doc2), ("sys3", sys3)):
        cubes = rng.random((16, PS.ndim))
        cubes[0, :] = 0.0
        cubes[1, :] = 1.0
        cubes[2, 0] = -1e-12           # out of the hypercube
        cubes[3, -1] = 1.0 + 1e-12     # out of the hypercube
        rows = []
        for x in cubes:
            inside = not (False in rutils.intervals(PS.hypercube, x))
            flat = [float(v) for v in rutils.cube_to_physical(PS, x)]
            rows.append({"cube": [float(v) for v in x], "inside": bool(inside), "physical": flat})
        plumb[name] = rows
    with open(os.path.join(HERE, "plumbing.json"), "w") as fh:
        json.dump(plumb, fh, indent=1)
    print("wrote systems.json known_answers.json plumbing.json")
    for k, v in systems.items():
        print(k, "ndim", v["ndim"], "t0/ftime/dt", v["t0"], v["ftime"], v["dt"],
              "nobs", [len(o["epochs"]) for o in v["observed"].values()])


def g5():
    """Regenerate the regression values of g5_long_baseline.json: the
    `trues` 3-planet solution (Examples/inputs/trues:13-23) against the full 1359-d tables, with
    the oracle. Needs only systems.json; run after a deliberate change of the canonical arithmetic
    and check that the two-decimal values of the survey's probe (78.56 / 44.08 / 30.38) still hold."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import oracle_py as o
    path = os.path.join(HERE, "g5_long_baseline.json")
    with open(path) as fh:
        g = json.load(fh)
    with open(os.path.join(HERE, "systems.json")) as fh:
        spec = dict(json.load(fh)[g["system"]])
    spec["mstar"], spec["rstarAU"] = g["mstar"], g["rstar"] * g["Rsun_to_AU"]
    osys = o.OracleSystem(spec)
    for dt, key in ((spec["dt"], "chi2_regression"), (0.1, "chi2_regression_dt0.1")):
        st, pl, ep, tm, rs, vs = o.ephemeris(g["flat_params"], spec["mstar"], spec["t0"], dt, spec["ftime"])
        tot, per = osys.chi2_from_events(pl, ep, tm, rs)
        assert st == 0 and [int((pl == i).sum()) for i in range(3)] == g["n_events"]
        assert all(abs(float(p) - q) <= 0.0051 for p, q in zip(per, g["chi2_survey_2dec"]))
        g[key] = [float(p) for p in per]
    with open(path, "w") as fh:
        json.dump(g, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "g5":
        g5()
    else:
        main()
