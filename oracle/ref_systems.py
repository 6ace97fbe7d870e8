"""The PlanetarySystem objects of the reference's notebooks, built with the reference's OWN classes
(SetPlanet / PlanetarySystem, planetarysystem.py:184-350) from its own observation tables:
  doc2   Documentation/2- Setup.ipynb cells 26-47   (t0=0, ftime=221, dt=0.18)
  sys3   Examples/simple_fit.ipynb cells 6-12       (t0=0, ftime=500)
  sys3f  same planets, default simulation()          (the full 1359-d tables)
`nau` is the imported reference package, `inputs` the directory holding Documentation/inputs and
Examples/inputs (the checkout, or baseline/_ref/inputs on the GPU box).
TEST / BENCH INFRASTRUCTURE (used by tests/ and by bench.py's reference-plumbing CPU leg)."""
import contextlib
import io
import os


def build(nau, inputs):
    d = os.path.join(inputs, "Documentation", "inputs")
    e = os.path.join(inputs, "Examples", "inputs")
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        Pb = nau.SetPlanet("Planet-b"); Pb.load_ttvs(os.path.join(d, "planetb_ttvs.dat"))
        Pb.mass = [1, 50.]; Pb.period = [5.35, 5.36]; Pb.ecc = [0., 0.2]; Pb.inclination = [90, 90]
        Pb.ascending_node = [88.36, 88.36]
        Pc = nau.SetPlanet("Planet-c"); Pc.load_ttvs(os.path.join(d, "planetc_ttvs.dat"))
        Pc.mass = [1, 50]; Pc.period = [11.83, 11.84]; Pc.ecc = [0, .2]; Pc.inclination = [90, 90]
        Pc.ascending_node = [70, 110.]
        doc2 = nau.PlanetarySystem("SystemX", mstar=0.9132, rstar=0.8632); doc2.add_planets([Pb, Pc])
        doc2.simulation(t0=0.0, ftime=221, dt=0.18)
        out["doc2"] = doc2
        for name, kw in (("sys3", dict(t0=0, ftime=500)), ("sys3f", {})):
            P1 = nau.SetPlanet('Planet-b'); P1.mass = [0.001, 10]; P1.period = [10.47, 10.48]; P1.ecc = [0.0, 0.1]
            P1.inclination = [90, 90]; P1.ascending_node = [60, 120]; P1.load_ttvs(os.path.join(e, "3pl_planet1_ttvs.dat"))
            P2 = nau.SetPlanet('Planet-c'); P2.mass = [0.001, 10]; P2.period = [14.10, 14.11]; P2.ecc = [0.0, 0.1]
            P2.inclination = [85, 95]; P2.ascending_node = [60, 120]; P2.load_ttvs(os.path.join(e, "3pl_planet2_ttvs.dat"))
            P3 = nau.SetPlanet('Planet-d'); P3.mass = [0.001, 10]; P3.period = [23.57, 23.58]; P3.ecc = [0.0, 0.1]
            P3.inclination = [90.3, 90.3]; P3.ascending_node = [88.5, 88.5]; P3.load_ttvs(os.path.join(e, "3pl_planet3_ttvs.dat"))
            s = nau.PlanetarySystem("MySystem", mstar=0.522, rstar=0.4422); s.add_planets([P1, P2, P3])
            s.simulation(**kw)
            out[name] = s
    return out


def with_spec(PS, spec):
    """Point a real PlanetarySystem at the simulation span and the observed table of `spec` (a
    system_spec dict): what bench.py's synthetic 1500-d workload needs. Only attributes that
    utils.log_likelihood_func reads are touched (planetarysystem.py:184-350)."""
    PS.t0, PS.ftime, PS.dt = spec["t0"], spec["ftime"], spec["dt"]
    PS.transit_times = {p: dict(zip(o["epochs"], o["times"])) for p, o in spec["observed"].items()}
    PS.sigma_obs = {p: dict(zip(o["epochs"], o["sigma"])) for p, o in spec["observed"].items()}
    PS.second_term_logL = {"sum": spec["second_term_sum"]}
    return PS
