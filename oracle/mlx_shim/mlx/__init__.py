"""TEST INFRASTRUCTURE ONLY -- a NumPy-backed stand-in for the slice of MLX 0.7.0 that the
reference's hot path touches.

MLX is not installable in this image (no network), so the reference's own files under
/root/reference cannot run as-is.  This shim lets `oracle/make_golden.py` import and execute the
reference's *unmodified* Python (sampling / embedding / NeRF.forward / raw2outputs / render_rays)
on the CPU in fp32, to produce the golden vectors in tests/golden/.  It pins the oracle's
restatement against the reference's code; it does NOT pin MLX-internal numerics (linspace formula,
sin/cos/exp ulps, reduction order), which are stated assumptions in DESIGN.md.

Never imported by the product package `nerf_meets_mlx_b200`.
"""
