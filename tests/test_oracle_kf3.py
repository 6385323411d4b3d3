"""3rd-order per-keypoint Kalman filter restatement (kalman_filter.cu) against numpy."""
import numpy as np


def test_initiate_predict_update_against_numpy(orc):
    rng = np.random.default_rng(5)
    T = 6
    dets = rng.uniform(0, 640, (T, 17, 3)).astype(np.float32)
    dets[..., 2] = rng.uniform(0, 1, (T, 17))
    dets[0, 3, 2] = 0.0
    k = orc.KF3(T)
    k.initiate(dets.reshape(T, 51), np.arange(T))
    m, d = k.state()
    m = m.reshape(T, 17, 8); d = d.reshape(T, 17, 8)
    assert np.array_equal(m[..., 0], dets[..., 0]) and np.array_equal(m[..., 1], dets[..., 1])
    assert (m[..., 2:] == 0).all()
    assert d[0, 3, 0] == 1000.0 and d[0, 0, 0] == 10.0 and (d[..., 2:] == 100.0).all()

    # give the state some velocity/acceleration by updating then predicting repeatedly
    f32 = np.float32
    for it in range(3):
        k.predict(T)
        m2 = m.copy(); d2 = d.copy()
        m2[..., 0] = m[..., 0] + m[..., 2] + f32(0.5) * m[..., 4] + f32(1.0 / 6.0) * m[..., 6]
        m2[..., 1] = m[..., 1] + m[..., 3] + f32(0.5) * m[..., 5] + f32(1.0 / 6.0) * m[..., 7]
        m2[..., 2] = m[..., 2] + m[..., 4] + f32(0.5) * m[..., 6]
        m2[..., 3] = m[..., 3] + m[..., 5] + f32(0.5) * m[..., 7]
        m2[..., 4:6] = m[..., 4:6] * f32(0.9); m2[..., 6:8] = m[..., 6:8] * f32(0.9)
        noise = np.array([1, 1, .5, .5, .1, .1, .05, .05], np.float32)
        d2 = d + noise * noise
        m, d = m2, d2
        gm, gd = k.state()
        assert np.array_equal(gm.reshape(T, 17, 8), m) and np.array_equal(gd.reshape(T, 17, 8), d)

        z = (dets + rng.normal(0, 3, dets.shape)).astype(np.float32)
        z[..., 2] = rng.uniform(0, 1, (T, 17)).astype(np.float32)
        matches = np.stack([np.arange(T), np.arange(T)[::-1]], 1)
        k.update(z.reshape(T, 51), matches)
        for t, di in matches:
            for kp in range(17):
                c = z[di, kp, 2]
                if c < f32(0.1):
                    continue
                yx = z[di, kp, 0] - m[t, kp, 0]; yy = z[di, kp, 1] - m[t, kp, 1]
                R = f32(5.0) / (c + f32(0.1))
                Kx = d[t, kp, 0] / (d[t, kp, 0] + R); Ky = d[t, kp, 1] / (d[t, kp, 1] + R)
                m[t, kp, 0] += Kx * yx; m[t, kp, 1] += Ky * yy
                Kv = f32(0.5) * Kx
                m[t, kp, 2] += Kv * yx; m[t, kp, 3] += Kv * yy
                d[t, kp, 0] = (f32(1.0) - Kx) * d[t, kp, 0]; d[t, kp, 1] = (f32(1.0) - Ky) * d[t, kp, 1]
        gm, gd = k.state()
        assert np.array_equal(gm.reshape(T, 17, 8), m) and np.array_equal(gd.reshape(T, 17, 8), d)

    out = k.extract(np.array([2, 0]))
    assert np.array_equal(out.reshape(2, 17, 3)[..., 0], m[[2, 0], :, 0]) and (out.reshape(2, 17, 3)[..., 2] == 1.0).all()
    mean, cov = k.full_state(1)
    assert np.array_equal(np.diag(cov), d[1].ravel()) and np.count_nonzero(cov - np.diag(np.diag(cov))) == 0
