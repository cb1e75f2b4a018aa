"""oracle/radiance_oracle.py (parity unpinned: the reference ships no vectors for its Slang kernels and slangtorch is
absent) checked against cases whose answer follows from the reference's formulas by hand."""
import numpy as np

from oracle import radiance_oracle as ro

IDENT = [1.0, 0.0, 0.0, 0.0]


def _pair(opacity_b=0.6, z=0.1, sh_dc=1.0):
    """Surfel A at the origin facing +z, surfel B at (0,0,z) facing -z (quaternion: rotation by pi about x)."""
    centers = np.array([[0, 0, 0], [0.02, -0.01, z]], np.float32)
    scales = np.full((2, 3), 0.05, np.float32)
    rot = np.array([IDENT, [0.0, 1.0, 0.0, 0.0]], np.float32)
    normals = np.array([[0, 0, 1], [0, 0, -1]], np.float32)
    opacity = np.array([0.9, opacity_b], np.float32)
    cov_inv = np.tile(np.array([1 / 0.05 ** 2, 0, 0, 1 / 0.05 ** 2, 0, 1 / 0.05 ** 2], np.float32), (2, 1))
    shs = np.zeros((2, 16, 3), np.float32)
    shs[:, 0] = sh_dc
    return ro.Surfels(centers, scales, rot, normals, opacity, cov_inv), shs


def test_rotation_matrix_of_the_identity_and_a_half_turn():
    m = ro.rotation_matrices(np.array([IDENT, [0.0, 1.0, 0.0, 0.0]], np.float32))
    np.testing.assert_allclose(m[0], np.eye(3), atol=1e-6)
    np.testing.assert_allclose(m[1], np.diag([1.0, -1.0, -1.0]), atol=1e-6)


def test_closest_hit_on_a_facing_surfel():
    sf, _ = _pair()
    i, t, alpha, uv = ro.closest_hit(sf, [0, 0, 0], [0, 0, 1], 0.042, 0.2)
    assert i == 1 and abs(t - 0.1) < 1e-6
    # the hit point is (-0.02, 0.01) from B's centre: power = -0.5 (0.02^2 + 0.01^2) / 0.05^2
    power = -0.5 * (0.02 ** 2 + 0.01 ** 2) / 0.05 ** 2
    assert abs(alpha - 0.6 * np.exp(power)) < 1e-6
    # local coordinates of the hit in B's frame (R = diag(1,-1,-1)): (-0.02, -0.01) / 0.05 -> swapped so that u >= v
    assert abs(uv[0] - (-0.2 * 0.5 + 0.5)) < 1e-6 and abs(uv[1] - (-0.4 * 0.5 + 0.5)) < 1e-6


def test_rejections():
    sf, _ = _pair()
    assert ro.closest_hit(sf, [0, 0, 0], [0, 0, 1], 0.11, 0.2)[0] == -1          # t < t_min
    assert ro.closest_hit(sf, [0, 0, 0], [0, 0, 1], 0.042, 0.1)[0] == -1         # t >= t_max (strictly closer only)
    assert ro.closest_hit(sf, [0, 0, 0.15], [0, 0, -1], 0.042, 0.2)[0] == 0      # from behind B: B's normal faces away, A is hit
    sf2, _ = _pair(opacity_b=0.003)
    assert ro.closest_hit(sf2, [0, 0, 0], [0, 0, 1], 0.042, 0.2)[0] == -1        # alpha < 1/255
    assert ro.closest_hit(sf, [0.3, 0, 0], [0, 0, 1], 0.042, 0.2)[0] == -1       # outside the 3-sigma ellipse


def test_radiance_visibility_and_first_hit_of_one_bounce():
    sf, shs = _pair(opacity_b=0.6, sh_dc=1.0)
    rad, vis, hit, uv = ro.render_radiance_with_sampling_SH(sf, shs, sf.c[:1], np.array([[[0, 0, 1], [0, 0, -1]]], np.float32))
    power = -0.5 * (0.02 ** 2 + 0.01 ** 2) / 0.05 ** 2
    alpha = 0.6 * np.exp(power)
    col = 0.28209479177387814 * 1.0 + 0.5
    np.testing.assert_allclose(rad[0, 0], col * alpha, rtol=1e-5)
    assert abs(vis[0, 0] - (1 - alpha)) < 1e-6 and hit[0, 0] == 1
    assert hit[0, 1] == -1 and vis[0, 1] == 1.0 and not rad[0, 1].any()           # nothing below
    # T < 0.2 after the hit -> visibility 0
    sf3, shs3 = _pair(opacity_b=0.99 / np.exp(power) + 1.0)
    _, vis3, _, _ = ro.render_radiance_with_sampling_SH(sf3, shs3, sf3.c[:1], np.array([[[0, 0, 1]]], np.float32))
    assert vis3[0, 0] == 0.0


def test_chunk_local_self_test_quirk():
    """self_mod > 0 ignores surfel (n % self_mod) -- here ray 1 of a 1-surfel chunk ignores surfel 0."""
    sf, shs = _pair()
    d = np.array([[[0, 0, -1]]], np.float32)                                       # from B down onto A
    _, _, hit_a, _ = ro.render_radiance_with_sampling_SH(sf, shs, sf.c[1:2] * [0, 0, 1], d, first_index=1, self_mod=0)
    _, _, hit_b, _ = ro.render_radiance_with_sampling_SH(sf, shs, sf.c[1:2] * [0, 0, 1], d, first_index=1, self_mod=1)
    assert hit_a[0, 0] == 0 and hit_b[0, 0] == -1


def test_brdf_and_direct_light_by_hand():
    n = np.array([0, 0, 1.0])
    v = ro.brdf_simple(n, n, n, np.array([0.5, 0.5, 0.5]), 1.0)
    # rough = 1: alpha2 = 1, k = 0.5, NoX = 1 -> nom = 4 pi, F = 0.04 + 0.96 * 2^(-12.53789)
    F = 0.04 + 0.96 * 2.0 ** (-5.55473 - 6.98316)
    np.testing.assert_allclose(v, F / (4 * np.pi) + 0.5 / np.pi, rtol=1e-12)
    env = np.arange(4 * 8 * 3, dtype=np.float64).reshape(4, 8, 3)
    # +x: theta = 0 -> column (We-1)/2, phi = pi/2 -> row (He-1)/2
    got = ro.direct_light(env, np.array([[1.0, 0, 0]]), 2.0)[0]
    want = 2.0 * 0.25 * (env[1, 3] + env[1, 4] + env[2, 3] + env[2, 4])
    np.testing.assert_allclose(got, want, rtol=1e-5)


def test_select_samples_takes_the_first_maximum():
    xyz = np.zeros((1, 3), np.float32); cam = np.array([0, 0, 1], np.float32); gn = np.array([[0, 0, 1]], np.float32)
    dirs = np.array([[[0, 0, 1], [0, 1, 0], [0, 0, 1]]], np.float32)
    # view = (0,0,-1); reflect = 2 (n.v) n + v = (0,0,-3); all visible -> all scores are +-0 -> index 0
    sel, _ = ro.select_samples(xyz, cam, gn, dirs, np.ones((1, 3), np.float32))
    assert sel[0] == 0
    sel, ndi = ro.select_samples(xyz, cam, gn, -dirs, np.array([[0.5, 0.0, 0.5]], np.float32))
    assert sel[0] == 0 and ndi[0, 0] == ndi[0, 2] == 1.5
