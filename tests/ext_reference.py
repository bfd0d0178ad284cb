"""fp64 restatement of the EXTENSION formulas (spherical SDEF, dual-quaternion QDEF, applied UV morphs).

PARITY UNPINNED: libmmd implements none of these, so there is no reference behaviour; this file
restates the formulas documented in include/mmdgpu.h / DESIGN.md independently of the CUDA code, in float64, and the
GPU tests compare within a tolerance.  Test infrastructure only."""
import numpy as np

from simple_mmd_renderer_b200 import capi


def app_slots(model):
    mt, mb, mc = model["morph_type"], model["morph_entry_begin"], model["morph_entry_count"]
    ge = model["group_morph_entries"]
    slots = []

    def visit(m, parent, mult):
        me = len(slots)
        slots.append((m, parent, float(np.float32(mult))))
        if mt[m] == capi.MORPH_GROUP:
            for j in range(int(mc[m])):
                e = ge[int(mb[m]) + j]
                visit(int(e["morph"]), me, e["rate"])
    for m in range(int(model["n_morphs"])):
        visit(m, -1, 1.0)
    return slots


def slot_rates(model, rates):
    """Application-slot rates as the device computes them: fp32 products, skipped when rate < 1e-7."""
    slots = app_slots(model)
    out = np.zeros(len(slots), np.float32)
    for s, (m, parent, mult) in enumerate(slots):
        if parent < 0:
            r = np.float32(rates[m])
        else:
            if out[parent] == 0:
                continue
            r = np.float32(np.float32(mult) * out[parent])
        out[s] = r if float(r) >= 1e-7 else 0.0
    return slots, out


def morph_images(model, rates):
    """(vertex displacement [nv,3], uv displacement [nv,2]) in float64."""
    slots, sr = slot_rates(model, rates)
    nv = int(model["n_vertices"])
    dv = np.zeros((nv, 3))
    duv = np.zeros((nv, 2))
    mt, mb, mc = model["morph_type"], model["morph_entry_begin"], model["morph_entry_count"]
    for s, (m, _, _) in enumerate(slots):
        r = float(sr[s])
        if r == 0.0:
            continue
        b0, n = int(mb[m]), int(mc[m])
        if mt[m] == capi.MORPH_VERTEX:
            e = model["vertex_morph_entries"][b0:b0 + n]
            np.add.at(dv, e["vertex"].astype(np.int64), e["offset"].astype(np.float64) * r)
        elif mt[m] == capi.MORPH_UV:
            e = model["uv_morph_entries"][b0:b0 + n]
            np.add.at(duv, e["vertex"].astype(np.int64), e["offset"][:, :2].astype(np.float64) * r)
    return dv, duv


def material_images(model, rates):
    """(n_materials, 2, 28) float64: multiplicative and additive material images, application order per material.
    MUL entry: mul *= 1 + (value - 1) * rate;  ADD entry: add += value * rate (include/mmdgpu.h)."""
    slots, sr = slot_rates(model, rates)
    nmat = int(model["n_materials"])
    img = np.zeros((nmat, 2, capi.MATERIAL_FIELDS))
    img[:, 0] = 1.0
    mt, mb, mc = model["morph_type"], model["morph_entry_begin"], model["morph_entry_count"]
    for s, (m, _, _) in enumerate(slots):
        r = float(sr[s])
        if r == 0.0 or mt[m] != capi.MORPH_MATERIAL:
            continue
        for e in model["material_morph_entries"][int(mb[m]):int(mb[m]) + int(mc[m])]:
            mi = int(e["material"])
            targets = range(nmat) if (mi < 0 or mi >= nmat) else [mi]
            v = e["value"].astype(np.float64)
            for t in targets:
                if int(e["method"]) == capi.MATERIAL_MUL:
                    img[t, 0] *= 1.0 + (v - 1.0) * r
                else:
                    img[t, 1] += v * r
    return img


def mat_to_quat(R):
    """Unit quaternion (x, y, z, w) of a column-vector rotation matrix R (v' = R v)."""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s]
    q = np.asarray(q, np.float64)
    return q / np.linalg.norm(q)


def qrot(q, v):
    u = q[:3]
    return v + 2.0 * np.cross(u, np.cross(u, v) + q[3] * v)


def bone_transforms(skin16):
    """Per bone: column-vector rotation R, translation t, quaternion q, dual part d from the row-vector 4x4."""
    out = []
    for M in skin16.reshape(-1, 4, 4).astype(np.float64):
        R = M[:3, :3].T
        t = M[3, :3]
        q = mat_to_quat(R)
        d = np.empty(4)
        d[:3] = 0.5 * (q[3] * t + np.cross(t, q[:3]))
        d[3] = -0.5 * np.dot(t, q[:3])
        out.append((R, t, q, d))
    return out


def sdef(p, n, b0, b1, w0, C, R0, R1, T):
    w1 = 1.0 - w0
    (Ra, ta, qa, _), (Rb, tb, qb, _) = T[b0], T[b1]
    dot = float(np.dot(qa, qb))
    if dot < 0:
        qb, dot = -qb, -dot
    if dot < 0.9995:
        om = np.arccos(dot)
        k0, k1 = np.sin(w0 * om) / np.sin(om), np.sin(w1 * om) / np.sin(om)
    else:
        k0, k1 = w0, w1
    q = qa * k0 + qb * k1
    q /= np.linalg.norm(q)
    rw = R0 * w0 + R1 * w1
    cr0 = (C + (C + R0 - rw)) * 0.5
    cr1 = (C + (C + R1 - rw)) * 0.5
    pos = qrot(q, p - C) + (Ra @ cr0 + ta) * w0 + (Rb @ cr1 + tb) * w1
    return pos, qrot(q, n)


def qdef(p, n, ids, w, T):
    qa = T[ids[0]][2]
    br, bd = np.zeros(4), np.zeros(4)
    for i in range(4):
        _, _, q, d = T[ids[i]]
        s = w[i] if np.dot(q, qa) >= 0 else -w[i]
        br += q * s
        bd += d * s
    nrm = np.linalg.norm(br)
    br, bd = br / nrm, bd / nrm
    t = 2.0 * (br[3] * bd[:3] - bd[3] * br[:3] + np.cross(br[:3], bd[:3]))
    return qrot(br, p) + t, qrot(br, n)
