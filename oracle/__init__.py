"""CPU oracle for the Scythe.jl semi-spectral hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy float64 restatement of the algorithm behind the
reference's hot path: the Springsteel-style spectral <-> grid
transforms (cubic B-spline radial, Fourier azimuthal, Chebyshev vertical), the
tile/patch decomposition, and Scythe.jl's time-step core.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the timed CPU
baseline -- never as something the product path (``scythe_jl_b200``) calls.

PARITY UNPINNED.  The reference (Scythe.jl v1.0.1) delegates every transform to
Springsteel.jl (``/root/reference/Project.toml:20``, uuid
66b52f22-0fad-4358-9559-2105d4560aaf, NO version pin, NO Manifest), whose source is
absent from ``/root/reference``; Julia itself is absent from this image.  The
reference has no tests and no golden vectors for this path.  The transform algebra
below therefore restates the *published* algorithm (Ooyama 2002 cubic-spline
transform; FFTW R2HC/HC2R and REDFT00 conventions) constrained by the reference's
call sites.  The only known answer in the tree -- the
LinearAdvection1D notebook output -- is reproduced to ~0.13 % (a sanity band, see
``tests/test_oracle_golden.py``), not to round-off.  The Scythe-level driver,
time-stepping and equation sets DO live in the reference and are restated with
file:line citations.
"""
