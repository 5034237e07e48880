"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU, fp32 restatement of the reference's (piljoong-jeong/nerf_meets_mlx) volume-learning hot
path, function by function, each citing the reference file:line it follows (paths relative to
/root/reference).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package; `nerf_meets_mlx_b200` never does.

Pinning status (DESIGN.md "Oracle"):
  * The reference ships no tests, golden vectors or fixtures for this path.
  * `oracle/make_golden.py` executes the reference's own, unmodified Python files from
    /root/reference under a NumPy-backed MLX stand-in (`oracle/mlx_shim`, because MLX 0.7.0 is not
    installable here) plus the real torch `sample_from_inverse_cdf_torch`, and commits the outputs
    as tests/golden/*.npz.  tests/test_oracle_golden.py checks this restatement against them.
  * MLX-internal numerics (linspace formula, transcendental ulps, reduction order, Adam without bias
    correction, tree-shared optimizer state) cannot be pinned without MLX: stated assumptions.
"""
