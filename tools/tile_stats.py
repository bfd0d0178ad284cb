#!/usr/bin/env python
"""Host-only statistics of the device vertex layout of a synthetic model: per 32-lane group, the shared-memory bank
conflict degree of the staging scatter (lanes whose PMX index is congruent mod 32 collide), the sliced-ELL padding, and
how many groups mix skinning types.  usage: python tools/tile_stats.py [C1|C2|C3] [binding]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dataclasses import replace
import numpy as np
from simple_mmd_renderer_b200 import capi, synth
from simple_mmd_renderer_b200.poser import plan_arrays

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = synth.CONFIGS[name]
if len(sys.argv) > 2:
    cfg = replace(cfg, binding=sys.argv[2])
model = synth.make_model(cfg)
t0 = time.time()
plan = plan_arrays(model)
t1 = time.time()
TILE, V = 512, 4
WARPS = TILE // V // 32
orig = plan[capi.PLAN_TILE_ORIG].astype(np.int64)
st = plan[capi.PLAN_TILE_TYPE].astype(np.int64)
n_tiles = orig.size // TILE
# storage position of (tile, group g = j * WARPS + w, lane l)
j, w, l = np.meshgrid(np.arange(V), np.arange(WARPS), np.arange(32), indexing="ij")
pos = ((w * 32 + l) * V + j).reshape(V * WARPS, 32)            # [group][lane]
o = orig.reshape(n_tiles, TILE)[:, pos]                        # [tile][group][lane]
ty = st.reshape(n_tiles, TILE)[:, pos]
res = o % 32
deg = np.zeros((n_tiles, V * WARPS), np.int64)
mixed = np.zeros((n_tiles, V * WARPS), bool)
for t in range(5):                                              # conflict domain = lanes of one type (divergent branches)
    m = ty == t
    if not m.any():
        continue
    cnt = np.zeros((n_tiles, V * WARPS, 32), np.int64)
    ti, gi, li = np.nonzero(m)
    np.add.at(cnt, (ti, gi, res[ti, gi, li]), 1)
    d = cnt.max(axis=2)
    present = m.sum(axis=2)
    # wavefronts of one STS of this branch = its worst bank; a group costs the sum over its branches
    deg += d
    mixed |= (present > 0) & (present < 32)
# palette loads: an LDS.128 takes 2 wavefronts instead of 4 when every aligned lane pair of the (type-uniform) instruction
# reads the same cell, i.e. agrees on the tile-local bone id at that position (profiles/r02_smem_patterns.txt)
lid = plan[capi.PLAN_TILE_LOCAL_ID].reshape(-1, 4).astype(np.int64).reshape(n_tiles, TILE, 4)[:, pos]   # [tile][group][lane][k]
keep_of = {0: 1, 1: 2, 2: 4, 3: 2, 4: 4}
wf_now = wf_all4 = 0.0
for t, keep in keep_of.items():
    m = ty == t
    if not m.any():
        continue
    present = m.any(axis=2)                                     # [tile][group]
    pair_same_type = m[:, :, 0::2] == m[:, :, 1::2]
    for k in range(keep):
        agree = (lid[:, :, 0::2, k] == lid[:, :, 1::2, k]) | ~m[:, :, 0::2]
        fast = (agree & pair_same_type).all(axis=2) & present
        wf_now += 6 * (2 * fast.sum() + 4 * (present & ~fast).sum())
        wf_all4 += 6 * 4 * present.sum()
print(f"  palette loads: {wf_now / n_tiles:.0f} wavefronts per tile and slot pair ({wf_all4 / n_tiles:.0f} with no pair on the 2-wavefront path)")
rounds = plan[capi.PLAN_ELL_ROUNDS].astype(np.int64)
row = plan[capi.PLAN_CSR_ROW_PTR].astype(np.int64)
real_entries = int(row[-1])
print(f"{cfg.name}: {n_tiles} tiles, plan built in {t1 - t0:.1f} s")
print(f"  staging scatter: {deg.mean():.3f} wavefronts per STS (1.0 = conflict-free); histogram {np.bincount(deg.ravel())[:8].tolist()}")
print(f"  groups with mixed types: {mixed.mean() * 100:.1f} %")
print(f"  sliced ELL: {rounds.sum() * 32} entries stored for {real_entries} real ones: padding {100.0 * (rounds.sum() * 32 / max(1, real_entries) - 1):.1f} %, mean rounds {rounds.mean():.3f}")
