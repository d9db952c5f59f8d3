"""shared set-up of the §8(f).4 tests: fill the caster lists of a PerFrameData up to the cbuffer's capacity"""
import numpy as np


def fill_casters(pf, n_point=5, n_spot=5, seed=0):
    """n_point point casters and n_spot spot casters over the synthetic height field (ranges chosen so that some pixels fall
    outside a light's range / frustum), each spot caster with its own shadow view matrix"""
    rng = np.random.default_rng(900 + seed)
    L = pf.Lights
    L.numPointCasters = n_point
    for i in range(n_point):
        l = L.point_casters[i]
        l.position.x, l.position.y, l.position.z = rng.uniform(-15, 15), rng.uniform(8, 16), rng.uniform(-15, 15)
        l.range = float(rng.uniform(18, 40))
        l.color.x, l.color.y, l.color.z = rng.uniform(0.4, 1.0, 3)
        l.brightness = float(rng.uniform(300, 900))
        l.depthBias = float(rng.uniform(0.0, 0.05))
    L.numSpotCasters = n_spot
    m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0; m[1] = 0.01; m[4] = -0.02; m[12] = 0.1
    for i in range(n_spot):
        s = L.spot_casters[i]
        s.position.x, s.position.y, s.position.z = rng.uniform(-12, 12), rng.uniform(9, 14), rng.uniform(-12, 12)
        s.spotDir.x, s.spotDir.y, s.spotDir.z = rng.uniform(-0.3, 0.3), -1.0, rng.uniform(-0.3, 0.3)
        s.innerConeAngle = float(rng.uniform(0.3, 0.5)); s.outerConeAngle = s.innerConeAngle + float(rng.uniform(0.15, 0.35))
        s.color.x, s.color.y, s.color.z = rng.uniform(0.4, 1.0, 3)
        s.brightness = float(rng.uniform(300, 900))
        s.depthBias = float(rng.uniform(0.0, 0.01))
        for k in range(16): L.shadowViews[i].m[k] = float(m[k]) * (1.0 + 0.2 * i)
    return pf
