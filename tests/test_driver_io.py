"""Driver-side pieces around the hot path: the reader of the reference's model files, the
launcher that replaces run_Scythe.jl, NetCDF output, and exact restarts through the launcher.  CPU: the kernels run in the
TEST-ONLY emulation build; the same checks run on the device in tests/test_gpu_parity.py::test_launcher_on_device."""
from pathlib import Path

import numpy as np
import pytest

import scythe_jl_b200 as S
from scythe_jl_b200 import modelfile, ncio, run as launcher

REF_MODELS = Path("/root/reference/models")

RL_MODEL = '''
# a two-layer slab configuration written the way the reference's model files are
model = ModelParameters(
    ts = 3.0,
    integration_time = 86400.0,   # one day
    output_interval = 120.0,
    equation_set = "Oneway_ShallowWater_Slab",
    initial_conditions = "./run/ic.csv",
    output_dir = "./run/",
    grid_params = GridParameters(
        geometry="RL",
        xmin = 0.0,
        xmax = 3.0e5,
        num_cells = 100,
        BCL = Dict(
            "h" => CubicBSpline.R1T1,
            "u" => CubicBSpline.R1T0),
        BCR = Dict(
            "h" => CubicBSpline.R0,
            "u" => CubicBSpline.R1T1),
        vars = Dict(
            "h" => 1,
            "u" => 2)),
    physical_params = Dict(
        :g => 9.81,
        :Cd => 2.4e-3,
	    :f => 5.0e-5),
    options = Dict(
        :semiimplicit => true))
'''


def test_model_file_reader_restricted_julia():
    m = modelfile.parse_model_text(RL_MODEL)
    assert isinstance(m, S.ModelParameters)
    assert (m.ts, m.integration_time, m.output_interval) == (3.0, 86400.0, 120.0)
    assert m.equation_set == "Oneway_ShallowWater_Slab" and m.output_dir == "./run/"
    gp = m.grid_params
    assert (gp.geometry, gp.xmin, gp.xmax, gp.num_cells) == ("RL", 0.0, 3.0e5, 100)
    assert gp.BCL == {"h": S.CubicBSpline.R1T1, "u": S.CubicBSpline.R1T0}
    assert gp.BCR == {"h": S.CubicBSpline.R0, "u": S.CubicBSpline.R1T1}
    assert gp.vars == {"h": 1, "u": 2} and isinstance(gp.num_cells, int)
    assert m.physical_params == {"g": 9.81, "Cd": 2.4e-3, "f": 5.0e-5}
    assert m.options == {"semiimplicit": True, "exact_reference_state": False}
    for bad, what in [("model = ModelParameters(ts = exp(1.0))", "unsupported constructor"),
                      ("model = ModelParameters(ts = 1.0", "expected"),
                      ("x = 3", "must assign"),
                      ("model = ModelParameters(ts = foo)", "unknown name"),
                      ("model = ModelParameters(ts = 1.0) @everywhere", "cannot read")]:
        with pytest.raises(modelfile.ModelFileError, match=what):
            modelfile.parse_model_text(bad)


@pytest.mark.skipif(not REF_MODELS.is_dir(), reason="the reference tree is only present in the build container")
def test_model_file_reader_on_the_reference_model_files():
    files = sorted(REF_MODELS.rglob("*.jl"))
    assert len(files) >= 4
    for f in files:
        m = modelfile.load_model_file(str(f))
        assert m.grid_params.num_cells == 100 and m.ts > 0 and m.equation_set
        names = m.grid_params.var_names()
        assert set(m.grid_params.BCL) == set(names) == set(m.grid_params.BCR)
        S.api._c_grid_params(m.grid_params)              # every BC constant resolves to a library code
    m = modelfile.load_model_file(str(REF_MODELS / "cha_bell2024" / "Twoway_ShallowWater_Slab.jl"))
    assert m.physical_params["S1"] == 1.0e-5 and m.grid_params.vars["wb"] == 6


def _rlz_params():
    return S.GridParameters(geometry="RLZ", xmin=0.0, xmax=10.0, num_cells=4, zmin=0.0, zmax=5.0, zDim=16,
                            BCL={"h": S.CubicBSpline.R1T1, "u": S.CubicBSpline.R1T0, "v": S.CubicBSpline.R1T0},
                            BCR={"h": S.CubicBSpline.R0, "u": S.CubicBSpline.R0, "v": S.CubicBSpline.R0},
                            vars={"h": 1, "u": 2, "v": 3})


def check_netcdf_round_trip(lib, tmp_path):
    from scipy.io import netcdf_file
    gp = _rlz_params()
    g = S.createGrid(gp, lib=lib)
    rng = np.random.default_rng(5)
    g.physical[:] = rng.standard_normal(g.physical.shape)
    want = g.physical.copy()
    path = str(tmp_path / "o" / "out.nc")
    ncio.write_grid_netcdf(g, path, 0.0, derivatives=True)
    g.physical[:, :, 0] *= 2.0
    ncio.write_grid_netcdf(g, path, 7.5, derivatives=True)
    f = netcdf_file(path, "r", mmap=False)
    assert f.Conventions == b"CF-1.8" and f.geometry == b"RLZ"
    hp = g.N // gp.zDim
    assert f.dimensions["point"] == hp and f.dimensions["z"] == gp.zDim and f.dimensions["time"] is None
    assert list(f.variables["time"][:]) == [0.0, 7.5]
    pts = S.getGridpoints(g).reshape(g.N, -1)
    assert np.array_equal(f.variables["r"][:], pts[::gp.zDim, 0]) and np.array_equal(f.variables["lambda"][:], pts[::gp.zDim, 1])
    assert np.array_equal(f.variables["z"][:], pts[:gp.zDim, 2])
    ring = f.variables["ring"][:]
    assert [int((ring == r).sum()) for r in range(3)] == [8, 12, 16]        # ring ri = 1, 2, 3 has 4 + 4 ri points
    for i, n in enumerate(("h", "u", "v")):
        assert f.variables[n].coordinates == b"r lambda"
        assert np.array_equal(f.variables[n][0].reshape(-1), want[:, i, 0])
        assert np.array_equal(f.variables[n][1].reshape(-1), 2.0 * want[:, i, 0])
        for d, s in enumerate(["r", "rr", "l", "ll", "z", "zz"], start=1):
            assert np.array_equal(f.variables[f"{n}_{s}"][0].reshape(-1), want[:, i, d])
    f.close()
    g.physical[:] = 0.0
    assert ncio.read_physical_grid_netcdf(path, g, record=0) == 0.0
    assert np.array_equal(g.physical[:, :, 0], want[:, :, 0])
    g.close()


def test_netcdf_round_trip(emu_lib, tmp_path):
    check_netcdf_round_trip(emu_lib, tmp_path)


def check_launcher(lib, tmp_path, monkeypatch, nsteps=6):
    """python -m scythe_jl_b200.run -w 2 model.jl: CSV + NetCDF output, a checkpoint on the way, and a second launch that
    restarts from it and ends on the same bytes."""
    from scipy.io import netcdf_file
    from scythe_jl_b200 import _lib
    monkeypatch.setattr(_lib, "_default", lib)
    gp = S.GridParameters(geometry="R", xmin=-50.0, xmax=50.0, num_cells=30, BCL={"u": S.CubicBSpline.PERIODIC},
                          BCR={"u": S.CubicBSpline.PERIODIC}, vars={"u": 1})
    g = S.createGrid(gp, lib=lib)
    x = S.getGridpoints(g)
    g.close()
    np.savetxt(tmp_path / "ic.csv", np.stack([x, np.exp(-(x / 20.0) ** 2)], 1), delimiter=",", header="r,u", comments="", fmt="%.17g")
    text = f'''model = ModelParameters(
        ts = 0.05, integration_time = {0.05 * nsteps}, output_interval = {0.05 * nsteps / 2},
        equation_set = "LinearAdvection1D",
        initial_conditions = "{tmp_path / "ic.csv"}", output_dir = "{tmp_path / "out"}/",
        grid_params = GridParameters(geometry = "R", xmin = -50.0, xmax = 50.0, num_cells = 30,
            BCL = Dict("u" => CubicBSpline.PERIODIC), BCR = Dict("u" => CubicBSpline.PERIODIC), vars = Dict("u" => 1)),
        physical_params = Dict(:c_0 => 1.0, :K => 0.0))'''
    (tmp_path / "model.jl").write_text(text)
    ck = str(tmp_path / "ck.npz")
    rc = launcher.main(["-w", "2", "--format", "both", "--checkpoint", ck, "--checkpoint-interval", str(0.05 * nsteps / 2),
                        str(tmp_path / "model.jl")])
    assert rc == 0
    tags = [str(round(t, 2)) for t in (0.0, 0.05 * nsteps / 2, 0.05 * nsteps)]
    assert sorted(p.name for p in (tmp_path / "out").iterdir()) == sorted([f"physical_out_{t}.csv" for t in tags] + ["scythe_out.nc"])
    final = np.loadtxt(tmp_path / "out" / f"physical_out_{tags[2]}.csv", delimiter=",", skiprows=1)[:, 1]
    f = netcdf_file(str(tmp_path / "out" / "scythe_out.nc"), "r", mmap=False)
    assert len(f.variables["time"][:]) == 3 and np.array_equal(f.variables["u"][2], final)
    f.close()
    # the checkpoint on disk is the one of the final step; rewind: run half, checkpoint, restart the second half
    (tmp_path / "model.jl").write_text(text.replace(f"integration_time = {0.05 * nsteps}", f"integration_time = {0.05 * nsteps / 2}")
                                       .replace(f'{tmp_path / "out"}/', f'{tmp_path / "half"}/'))
    assert launcher.main(["-w", "2", "--checkpoint", ck, str(tmp_path / "model.jl")]) == 0
    (tmp_path / "model.jl").write_text(text.replace(f'{tmp_path / "out"}/', f'{tmp_path / "rest"}/'))
    assert launcher.main(["-w", "2", "--restart", ck, str(tmp_path / "model.jl")]) == 0
    again = np.loadtxt(tmp_path / "rest" / f"physical_out_{tags[2]}.csv", delimiter=",", skiprows=1)[:, 1]
    assert np.array_equal(again, final)
    assert launcher.main(["--sge", str(tmp_path / "model.jl")]) == 2
    assert launcher.main(["--gpus", "2", "-w", "3", str(tmp_path / "model.jl")]) == 2


def test_launcher_csv_netcdf_checkpoint_restart(emu_lib, tmp_path, monkeypatch):
    check_launcher(emu_lib, tmp_path, monkeypatch)
