"""shared set-up of the §8(f).1 surface-producer tests: synthetic materials, their mip chains (oracle-built), inputs"""
import numpy as np

import oracle_lib as orc
import vqengine_b200 as vq
from vqengine_b200 import synth


def material_set(n=4, tex_res=64, seed=synth.SEED_BASE + 11, uniform=False):
    mats, texs = synth.materials(n, tex_res, seed, uniform)
    chains = []
    for t in texs:
        d = {}
        for slot, lvl0 in t.items():
            if lvl0 is None:
                d[slot] = None
            else:
                h, w = lvl0.shape[:2]
                levels = vq.mip_level_count(w, h)
                d[slot] = (orc.texture_mip_chain(lvl0, levels), w, h, levels)
        chains.append(d)
    return mats, texs, chains
