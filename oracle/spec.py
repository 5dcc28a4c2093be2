"""Constants of the splatting hot path, shared by the oracle stages.

TEST INFRASTRUCTURE ONLY. Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker.

PARITY UNPINNED UPSTREAM: the reference tree (/root/reference) holds no tests,
fixtures or golden vectors for this path, and the CUDA sources it pip-installs
(ashawkey/diff-gaussian-rasterization, DSaurus/simple-knn, both unpinned in
README.md:17-20) are not vendored.  The constants below restate the published
algorithm of those packages and are anchored on the
reference's own call sites:

  * tile size / channel count        renderer/diff_gaussian_rasterizer.py:83-131 (3-channel images)
  * SH basis constants               geometry/sugar.py:744-772 (identical basis and sign convention)
  * colour = clamp_min(sh + 0.5, 0)  geometry/sugar.py:668-669
  * znear / zfar / fovx := fovy      renderer/gaussian_batch_renderer.py:24-26
"""

BLOCK_X = 16
BLOCK_Y = 16
BLOCK_SIZE = BLOCK_X * BLOCK_Y
NUM_CHANNELS = 3

NEAR_CULL = 0.2          # p_view.z <= 0.2 -> culled
FOV_CLAMP = 1.3          # t.xy / t.z clamped to +-1.3 tan(fov/2)
DILATION = 0.3           # low-pass added to cov2D diagonal
LAMBDA_FLOOR = 0.1       # max(0.1, mid^2 - det)
RADIUS_SIGMAS = 3.0      # radius = ceil(3 sqrt(lambda_max))
ALPHA_MAX = 0.99
ALPHA_MIN = 1.0 / 255.0
T_MIN = 1.0e-4
PW_EPS = 1.0e-7          # p_w = 1 / (p_hom.w + 1e-7)
DENOM2_EPS = 1.0e-7      # backward conic: 1 / (det^2 + 1e-7)

# geometry/sugar.py:745-764
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (
    1.0925484305920792,
    -1.0925484305920792,
    0.31539156525252005,
    -1.0925484305920792,
    0.5462742152960396,
)
SH_C3 = (
    -0.5900435899266435,
    2.890611442640554,
    -0.4570457994644658,
    0.3731763325901154,
    -0.4570457994644658,
    1.445305721320277,
    -0.5900435899266435,
)

KNN_K = 3


def higher_msb(n: int) -> int:
    """Upstream ``getHigherMsb``: binary search for the bit above the MSB of n.

    Returns 7/9/11/13 for n = 64/256/1024/4096 tiles.
    """
    msb = 32 // 2
    step = msb
    while step > 1:
        step //= 2
        if n >> msb:
            msb += step
        else:
            msb -= step
    if n >> msb:
        msb += 1
    return msb
