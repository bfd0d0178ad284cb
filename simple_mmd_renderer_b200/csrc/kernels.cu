// kernels.cu — the sm_100a kernels of the deformation path.
//
//   K1 pose_sample_kernel   VMD keyframe lookup + Bezier table lerp + NLerp      (Motion::GetBonePose/GetMorphPose)
//   K2 hierarchy_kernel     morph rates, bone morphs, bone program in waves, CCD IK, skinning palette
//   K3 skin_kernel          vertex morph gather + BDEF1/2/4 skinning of positions and normals (Poser::Deform)
//
// Compiled with -fmad=false: the fp32 expression trees are libmmd's, un-contracted.
#include <cstdint>
#include <cstdlib>

#include "device_types.cuh"
#include "kernels.cuh"
#include "mmd_math.cuh"

namespace mmdgpu {

using namespace dm;

// Programmatic dependent launch (one-slot objects: the interactive frame loop is three small dependent kernels).  A kernel
// launched with the attribute may start while its predecessor in the stream still runs; everything the predecessor wrote is
// only visible after pdl_wait(), so the model's static data is fetched before it and the slot state after it.  Both are
// no-ops in a launch without the attribute / without a dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// =================================================================================================
// K1 — keyframe sampling.  One thread per (slot, bone) and (slot, morph).
// Motion::GetBonePose(name, size_t)  L/motion/motion_impl.inl:255-319
// Motion::GetMorphPose(name, size_t) L/motion/motion_impl.inl:382-424
// MotionPlayer::SeekFrame            L/motion/poser_impl.inl:539-546
// Poser::ResetPosing (pose part)     L/motion/poser_impl.inl:131-137   (write_untracked = 1)
// =================================================================================================
// upper_bound for a key strictly inside (a[0], a[n-1]): four probes around the position a uniform key spacing predicts are
// loaded together and narrow [lo, hi) before the bisection, which then usually has nothing left to do - the search is a
// chain of dependent loads (log2 n of them) on the latency-bound interactive path.  Returns the first index whose key is greater (std::upper_bound).
__device__ __forceinline__ uint32_t upper_bound_guess_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t key, uint32_t first,
                                                          uint32_t last) {
    uint32_t lo = 1, hi = n - 1;   // a[0] <= key < a[n-1] is known
    const uint32_t g = (uint32_t)(((unsigned long long)(key - first) * (n - 1)) / (last - first));   // 0 .. n-2
    const uint32_t p0 = g > 0 ? g - 1 : 0, p1 = g, p2 = min(g + 1, n - 1), p3 = min(g + 2, n - 1);
    const uint32_t v0 = __ldg(a + p0), v1 = __ldg(a + p1), v2 = __ldg(a + p2), v3 = __ldg(a + p3);
    if (v0 <= key) lo = max(lo, p0 + 1); else hi = min(hi, p0);
    if (v1 <= key) lo = max(lo, p1 + 1); else hi = min(hi, p1);
    if (v2 <= key) lo = max(lo, p2 + 1); else hi = min(hi, p2);
    if (v3 <= key) lo = max(lo, p3 + 1); else hi = min(hi, p3);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Bracketing of one track at one slot: either one key verbatim (`use`) or keys (l, r) with barycentre `bary`.
//   frame mode  Motion::GetBonePose(name, size_t), motion_impl.inl:266-291: `left.frame == frame` returns the left key
//   time mode   Motion::GetBonePose(name, double), motion_impl.inl:333-354: frame = seconds*30 as a double, bracket by
//               upper_bound(size_t(frame)), no equality shortcut, barycentre computed in double then cast to float
struct Bracket { uint32_t use, l, r; float bary; };
__device__ __forceinline__ Bracket bracket_keys(const uint32_t* __restrict__ kf, uint32_t n, bool time_mode, uint32_t frame,
                                                double dframe) {
    Bracket b{0xFFFFFFFFu, 0u, 0u, 0.0f};
    const uint32_t first = __ldg(kf), last = __ldg(kf + n - 1);
    if (!time_mode) {
        if (first >= frame) b.use = 0;
        else if (last <= frame) b.use = n - 1;
        else {
            b.r = upper_bound_guess_u32(kf, n, frame, first, last);
            b.l = b.r - 1;
            const uint32_t lf = kf[b.l], rf = kf[b.r];
            if (lf == frame) b.use = b.l;
            else b.bary = (float)(frame - lf) / (float)(rf - lf);
        }
    } else {
        if ((double)first >= dframe) b.use = 0;
        else if ((double)last <= dframe) b.use = n - 1;
        else {
            b.r = upper_bound_guess_u32(kf, n, (uint32_t)(unsigned long long)dframe, first, last);
            b.l = b.r - 1;
            const uint32_t lf = kf[b.l], rf = kf[b.r];
            b.bary = (float)((dframe - (double)lf) / (double)(rf - lf));
        }
    }
    return b;
}

// How the slots of a launch get their frame.  by_value: the (single) frame id / time arrives as a kernel argument instead
// of through F.frame_id / F.time_s - the interactive path (one Poser, one frame per call) then needs no host-to-device copy.
struct SampleArgs {
    const DevAnim* anims;      // [n_instances], nullptr: ResetPosing (identity / zero everywhere)
    uint32_t write_untracked;  // also write identity / zero for bones and morphs the clip does not animate (ResetPosing + SeekFrame)
    uint32_t range_mode, frame_stride, time_mode, by_value, frame0;
    double time0;
    uint32_t n_inline;         // > 0: frame ids come from `inline_ids` instead of F.frame_id
    uint32_t inline_ids[kInlineFrameIds];
};
struct SlotFrame { uint32_t inst, frame; double dframe; };
__device__ __forceinline__ SlotFrame frame_of_slot(const DevFrames& F, const SampleArgs& A, uint32_t slot) {
    SlotFrame s{slot / F.n_frames, 0u, 0.0};
    if (A.time_mode) s.dframe = (A.by_value ? A.time0 : F.time_s[slot]) * 30.0;
    else if (A.by_value) s.frame = A.frame0 + (A.range_mode ? (slot - s.inst * F.n_frames) * A.frame_stride : 0u);
    else if (A.n_inline) s.frame = A.range_mode ? (A.inline_ids[s.inst] + (slot - s.inst * F.n_frames) * A.frame_stride) : A.inline_ids[slot];
    else s.frame = A.range_mode ? (F.frame_id[s.inst] + (slot - s.inst * F.n_frames) * A.frame_stride) : F.frame_id[slot];
    return s;
}
// Motion::GetBonePose for model bone b; returns false if the pose in the Poser must be left as it is (bone not in the clip
// and no reset requested)
__device__ __forceinline__ bool sample_bone(const SampleArgs& SA, const SlotFrame& sf, uint32_t b, float4& T, float4& R) {
    T = make_float4(0.f, 0.f, 0.f, 0.f); R = make_float4(0.f, 0.f, 0.f, 1.f);
    bool tracked = false;
    if (SA.anims) {
        const DevAnim A = SA.anims[sf.inst];
        tracked = A.bone_tracked[b] != 0;
        const uint32_t n = A.bone_key_count[b];
        if (tracked && n > 0) {
            const uint32_t k0 = A.bone_key_begin[b];
            const Bracket br = bracket_keys(A.key_frame + k0, n, SA.time_mode != 0, sf.frame, sf.dframe);
            if (br.use != 0xFFFFFFFFu) {
                T = A.key_T[k0 + br.use];
                R = A.key_R[k0 + br.use];
            } else {
                const float bary = br.bary;
                const float4 lT = A.key_T[k0 + br.l], rT = A.key_T[k0 + br.r];
                const float4 lR = A.key_R[k0 + br.l], rR = A.key_R[k0 + br.r];
                const uint4 cv = A.key_curve[k0 + br.l];  // curves of the LEFT key (motion_impl.inl:302-312)
                float lam = bezier_at(A.tables, cv.x, bary);
                T.x = lT.x * (1 - lam) + rT.x * lam;
                lam = bezier_at(A.tables, cv.y, bary);
                T.y = lT.y * (1 - lam) + rT.y * lam;
                lam = bezier_at(A.tables, cv.z, bary);
                T.z = lT.z * (1 - lam) + rT.z * lam;
                const float l_ = bezier_at(A.tables, cv.w, bary);
                R = v4_nlerp(lR, rR, l_);
            }
            T.w = 0.f;
        }
    }
    return tracked || SA.write_untracked;
}
// Motion::GetMorphPose for model morph m
__device__ __forceinline__ bool sample_morph(const SampleArgs& SA, const SlotFrame& sf, uint32_t m, float& w) {
    w = 0.0f;
    bool tracked = false;
    if (SA.anims) {
        const DevAnim A = SA.anims[sf.inst];
        tracked = A.morph_tracked[m] != 0;
        const uint32_t n = A.morph_key_count[m];
        if (tracked && n > 0) {
            const uint32_t k0 = A.morph_key_begin[m];
            const float* kw = A.mkey_weight + k0;
            const Bracket br = bracket_keys(A.mkey_frame + k0, n, SA.time_mode != 0, sf.frame, sf.dframe);
            if (br.use != 0xFFFFFFFFu) w = kw[br.use];
            else {
                const float lam = br.bary;  // default-constructed Bezier is linear (math_impl.inl:1350-1354)
                w = kw[br.l] * (1 - lam) + kw[br.r] * lam;
            }
        }
    }
    return tracked || SA.write_untracked;
}

__global__ void __launch_bounds__(128) pose_sample_kernel(DevModel M, DevFrames F, SampleArgs SA) {
    pdl_trigger();
    const uint32_t slot = blockIdx.y;
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (!SA.anims && F.material_images && blockIdx.x == 0) {
        // ResetPosing: all rates are zero, so the material images (extension) are their initial 1 / 0
        const uint32_t n = M.n_materials * 2 * MMDGPU_MATERIAL_FIELDS;
        float* mi = F.material_images + (size_t)slot * n;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mi[i] = ((i / MMDGPU_MATERIAL_FIELDS) & 1u) ? 0.0f : 1.0f;
    }
    if (item >= M.nb + M.nm) return;
    const SlotFrame sf = frame_of_slot(F, SA, slot);
    if (item < M.nb) {
        float4 T, R;
        if (sample_bone(SA, sf, item, T, R)) {
            F.poseT[(size_t)slot * M.nb + item] = T;
            F.poseR[(size_t)slot * M.nb + item] = R;
        }
    } else {
        const uint32_t m = item - M.nb;
        float w;
        if (sample_morph(SA, sf, m, w)) F.rate[(size_t)slot * M.nm + m] = w;
    }
}

// =================================================================================================
// K2 — hierarchy.  One warp per slot; lanes execute the ops of one wave in parallel.
// =================================================================================================
struct SlotState {
    const BoneStatic* bones;  // static per-bone records (global memory, or the CTA's shared-memory copy)
    const float4* poseR;
    const float4* poseT;
    float4* totR;
    float4* totT;
    float4* local;   // 3 float4 per bone
    float4* ikR;
    float4* preIK;
    const float4* morphR;
    const float4* morphT;
    float4* palette; // 3 float4 per bone
    float4* pal_ext; // extensions: 2 float4 per bone (rotation quaternion, dual part), else nullptr
};

__device__ __forceinline__ Mat43 load_local(const float4* p) {
    const float4 a = p[0], b = p[1], c = p[2];
    Mat43 M;
    M.m[0][0] = a.x; M.m[0][1] = a.y; M.m[0][2] = a.z; M.m[1][0] = a.w;
    M.m[1][1] = b.x; M.m[1][2] = b.y; M.m[2][0] = b.z; M.m[2][1] = b.w;
    M.m[2][2] = c.x; M.m[3][0] = c.y; M.m[3][1] = c.z; M.m[3][2] = c.w;
    return M;
}
__device__ __forceinline__ void store_local(float4* p, const Mat43& M) {
    p[0] = make_float4(M.m[0][0], M.m[0][1], M.m[0][2], M.m[1][0]);
    p[1] = make_float4(M.m[1][1], M.m[1][2], M.m[2][0], M.m[2][1]);
    p[2] = make_float4(M.m[2][2], M.m[3][0], M.m[3][1], M.m[3][2]);
}
__device__ __forceinline__ BoneStatic load_bone(const BoneStatic* __restrict__ bones, int32_t b) {
    const float4* p = reinterpret_cast<const float4*>(bones + b);
    const float4 a = p[0], c = p[1], d = p[2];
    BoneStatic s;
    s.local_offset[0] = a.x; s.local_offset[1] = a.y; s.local_offset[2] = a.z; s.parent = __float_as_int(a.w);
    s.position[0] = c.x; s.position[1] = c.y; s.position[2] = c.z; s.append_parent = __float_as_int(c.w);
    s.append_ratio = d.x; s.flags = (uint32_t)__float_as_int(d.y);
    s.link_slot = __float_as_int(d.z); s.morph_slot = __float_as_int(d.w);
    return s;
}

// local_matrix_ from total rotation / translation, then * parent (poser_impl.inl:161-166, :294-299)
__device__ __forceinline__ void set_local(const SlotState& S, const BoneStatic& s, int32_t b, const Quat& totR,
                                          const float4& totT) {
    Mat43 L;
    q_to_rows(totR, L);
    L.m[3][0] = totT.x + s.local_offset[0];
    L.m[3][1] = totT.y + s.local_offset[1];
    L.m[3][2] = totT.z + s.local_offset[2];
    if (s.flags & kHasParent) {
        // a bone that is its own parent multiplies the matrix it has just built by itself: libmmd's
        // local_matrix_ = local_matrix_ * bone_images_[parent_].local_matrix_ names one object twice
        const Mat43 P = (s.parent == b) ? L : load_local(S.local + 3 * (size_t)s.parent);
        L = m_mul(L, P);
    }
    store_local(S.local + 3 * (size_t)b, L);
}

// Poser::UpdateBoneTransform(size_t) without the IK part, L/motion/poser_impl.inl:142-166
__device__ __forceinline__ void eval_bone(const SlotState& S, int32_t b) {
    const BoneStatic s = load_bone(S.bones, b);
    const Quat R = q_from(S.poseR[b]);
    const float4 T = S.poseT[b];
    Quat mR = q_identity();
    float4 mT = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s.morph_slot >= 0) {
        mR = q_from(S.morphR[s.morph_slot]);
        mT = S.morphT[s.morph_slot];
    }
    Quat totR = q_mul(mR, R);
    float4 totT = make_float4(mT.x + T.x, mT.y + T.y, mT.z + T.z, 0.f);
    if (s.flags & (kAppendRot | kAppendTrans)) {
        // libmmd updates total_rotation_ / total_translation_ in place (poser_impl.inl:144-156): a bone that names
        // ITSELF as its append parent reads the values just written, not last frame's / the reset ones
        const bool self = s.append_parent == b;
        if (s.flags & kAppendRot) {
            const Quat pr = self ? totR : q_from(S.totR[s.append_parent]);
            totR = q_mul(totR, q_slerp(q_identity(), pr, s.append_ratio));
        }
        if (s.flags & kAppendTrans) {
            const float4 pt = self ? totT : S.totT[s.append_parent];
            totT.x = totT.x + s.append_ratio * pt.x;
            totT.y = totT.y + s.append_ratio * pt.y;
            totT.z = totT.z + s.append_ratio * pt.z;
        }
    }
    if (s.flags & kIsLink) {
        S.preIK[s.link_slot] = q_to4(totR);
        totR = q_mul(q_from(S.ikR[s.link_slot]), totR);
    }
    S.totR[b] = q_to4(totR);
    S.totT[b] = totT;
    set_local(S, s, b, totR, totT);
}
__device__ __forceinline__ void eval_bone(const DevModel&, const SlotState& S, int32_t b) { eval_bone(S, b); }

__device__ __forceinline__ Vec3 local_pos(const SlotState& S, int32_t b) {
    const float4 c = S.local[3 * (size_t)b + 2];
    return Vec3{c.y, c.z, c.w};
}

// CCD IK, L/motion/poser_impl.inl:168-310 (the part of UpdateBoneTransform after the bone's own transform).
//
// Nested solves: libmmd re-evaluates the links and the target with UpdateBoneTransform itself (:203-206, :303), so a
// link or target that has IK runs ITS solve at that point - for the target, after every CCD step.  DEPTH counts the
// levels above this one; the host refuses models nested deeper than kMaxIkDepth.  A nested solve is a real call (one
// copy of the code per level, cold), the top level stays inlined in the kernels.
struct IkTables {
    const IkDesc* iks;
    const IkLink* links;
};
template <int DEPTH>
__device__ __noinline__ void solve_ik_nested(IkTables T, SlotState S, uint32_t ik_index);

// UpdateBoneTransform(b) as the solver calls it: the bone's own transform, then its solve if it has one
// NEST = false is the kernel every model without nested solves runs: no call, no extra registers.
template <int DEPTH, bool NEST>
__device__ __forceinline__ void eval_bone_and_solve(const IkTables& T, const SlotState& S, int32_t b) {
    eval_bone(S, b);
    if (NEST && DEPTH + 1 < kMaxIkDepth) {
        const uint32_t flags = S.bones[b].flags;
        if (flags & kHasIk) solve_ik_nested<DEPTH + 1>(T, S, flags >> 16);
    }
}

template <int DEPTH, bool NEST>
__device__ __forceinline__ void solve_ik(const IkTables& T, const SlotState& S, const IkDesc k) {
    const IkLink* __restrict__ links = T.links + k.link_begin;
    const int nl = k.link_count;
    for (int i = 0; i < nl; ++i) S.ikR[S.bones[links[i].bone].link_slot] = make_float4(0.f, 0.f, 0.f, 1.f);
    const Vec3 ik_pos = local_pos(S, k.bone);
    for (int i = 0; i < nl; ++i) eval_bone_and_solve<DEPTH, NEST>(T, S, links[nl - i - 1].bone);
    eval_bone_and_solve<DEPTH, NEST>(T, S, k.target);
    Vec3 tp = local_pos(S, k.target);
    Vec3 err{ik_pos.x - tp.x, ik_pos.y - tp.y, ik_pos.z - tp.z};
    if ((double)v_dot(err, err) < kEpsD) return;
    const int iters = k.iterations;
    const int ikt = iters / 2;
    for (int i = 0; i < iters; ++i) {
        for (int j = 0; j < nl; ++j) {
            const IkLink L = links[j];
            if (L.fix == 4) continue;
            const BoneStatic ls = load_bone(S.bones, L.bone);
            const Vec3 lp = local_pos(S, L.bone);
            Vec3 td = v_normalize(Vec3{lp.x - tp.x, lp.y - tp.y, lp.z - tp.z});
            Vec3 id = v_normalize(Vec3{lp.x - ik_pos.x, lp.y - ik_pos.y, lp.z - ik_pos.z});
            // Triple::operator*, L/util/math_impl.inl:260-266
            Vec3 ax{td.y * id.z - td.z * id.y, td.z * id.x - td.x * id.z, td.x * id.y - td.y * id.x};
            if ((double)fabsf(ax.x) < kEpsD) ax.x = (float)kEpsD;
            if ((double)fabsf(ax.y) < kEpsD) ax.y = (float)kEpsD;
            if ((double)fabsf(ax.z) < kEpsD) ax.z = (float)kEpsD;
            Mat43 P = m_identity();
            if (ls.flags & kHasParent) P = load_local(S.local + 3 * (size_t)ls.parent);
            if (L.limited && L.fix != 0 && i < ikt) {
                const int r = L.fix - 1;
                const float d = ax.x * P.m[r][0] + ax.y * P.m[r][1] + ax.z * P.m[r][2];
                const float sgn = (d >= 0.0f) ? 1.0f : -1.0f;
                ax = Vec3{r == 0 ? sgn : 0.0f, r == 1 ? sgn : 0.0f, r == 2 ? sgn : 0.0f};
            } else {
                // rotate(axis, P^T), L/util/math_impl.inl:1032-1038
                const Vec3 t{ax.x * P.m[0][0] + ax.y * P.m[0][1] + ax.z * P.m[0][2],
                             ax.x * P.m[1][0] + ax.y * P.m[1][1] + ax.z * P.m[1][2],
                             ax.x * P.m[2][0] + ax.y * P.m[2][1] + ax.z * P.m[2][2]};
                ax = v_normalize(t);
            }
            const float ang = s_min(m_acos(m_clamp(v_dot(td, id), -1.0f, 1.0f)), k.angle_limit * (float)(j + 1));
            Quat ikR = q_mul(axis_to_quat(ax, ang), q_from(S.ikR[ls.link_slot]));
            if (L.limited) {
                const Quat pre = q_from(S.preIK[ls.link_slot]);
                Quat lr = q_mul(ikR, pre);
                Vec3 eu = quat_to_euler(L.order, lr);
                const bool refl = i < ikt;
                eu.x = limit_one(eu.x, L.lo[0], L.hi[0], refl);
                eu.y = limit_one(eu.y, L.lo[1], L.hi[1], refl);
                eu.z = limit_one(eu.z, L.lo[2], L.hi[2], refl);
                lr = euler_to_quat(L.order, eu);
                ikR = q_mul(lr, q_inverse(pre));
            }
            S.ikR[ls.link_slot] = q_to4(ikR);
            for (int c = j; c >= 0; --c) {  // links j .. 0 (poser_impl.inl:292-300)
                const int32_t cb = links[c].bone;
                const BoneStatic cs = load_bone(S.bones, cb);
                const Quat tot = q_mul(q_from(S.ikR[cs.link_slot]), q_from(S.preIK[cs.link_slot]));
                S.totR[cb] = q_to4(tot);
                set_local(S, cs, cb, tot, S.totT[cb]);
            }
            eval_bone_and_solve<DEPTH, NEST>(T, S, k.target);
            tp = local_pos(S, k.target);
        }
        err = Vec3{ik_pos.x - tp.x, ik_pos.y - tp.y, ik_pos.z - tp.z};
        if (v_dot(err, err) < kEpsF) return;
    }
}

template <int DEPTH>
__device__ __noinline__ void solve_ik_nested(IkTables T, SlotState S, uint32_t ik_index) {
    if (DEPTH < kMaxIkDepth) solve_ik<(DEPTH < kMaxIkDepth ? DEPTH : kMaxIkDepth - 1), true>(T, S, T.iks[ik_index]);
}

// Poser::UpdateBoneSkinningMatrix, L/motion/poser_impl.inl:320-326: skin = global_offset * local, stored as
// the three columns the skinning kernel consumes.
__device__ __forceinline__ void skin_bone(const DevModel& M, const SlotState& S, int32_t b) {
    const BoneStatic s = load_bone(S.bones, b);
    const Mat43 L = load_local(S.local + 3 * (size_t)b);
    const float g30 = -s.position[0], g31 = -s.position[1], g32 = -s.position[2];
    float4 col[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // rows of global_offset are unit vectors; the zero products are kept for signed-zero fidelity
        const float r0 = 1.0f * L.m[0][c] + 0.0f * L.m[1][c] + 0.0f * L.m[2][c] + 0.0f * L.m[3][c];
        const float r1 = 0.0f * L.m[0][c] + 1.0f * L.m[1][c] + 0.0f * L.m[2][c] + 0.0f * L.m[3][c];
        const float r2 = 0.0f * L.m[0][c] + 0.0f * L.m[1][c] + 1.0f * L.m[2][c] + 0.0f * L.m[3][c];
        const float r3 = g30 * L.m[0][c] + g31 * L.m[1][c] + g32 * L.m[2][c] + 1.0f * L.m[3][c];
        col[c] = make_float4(r0, r1, r2, r3);
    }
    S.palette[3 * (size_t)b + 0] = col[0];
    S.palette[3 * (size_t)b + 1] = col[1];
    S.palette[3 * (size_t)b + 2] = col[2];
    if (S.pal_ext) {
        // Extensions (no libmmd counterpart): the rotation of the skinning transform as a unit quaternion q and the
        // dual part d = 0.5 * (t, 0) (x) q for dual-quaternion blending.  Column-vector rotation R[i][j] = col[i][j].
        const float r00 = col[0].x, r01 = col[0].y, r02 = col[0].z, r10 = col[1].x, r11 = col[1].y, r12 = col[1].z,
                    r20 = col[2].x, r21 = col[2].y, r22 = col[2].z;
        float qx, qy, qz, qw;
        const float tr = r00 + r11 + r22;
        if (tr > 0.0f) {
            const float s4 = sqrtf(tr + 1.0f) * 2.0f;
            qw = 0.25f * s4; qx = (r21 - r12) / s4; qy = (r02 - r20) / s4; qz = (r10 - r01) / s4;
        } else if (r00 > r11 && r00 > r22) {
            const float s4 = sqrtf(1.0f + r00 - r11 - r22) * 2.0f;
            qw = (r21 - r12) / s4; qx = 0.25f * s4; qy = (r01 + r10) / s4; qz = (r02 + r20) / s4;
        } else if (r11 > r22) {
            const float s4 = sqrtf(1.0f + r11 - r00 - r22) * 2.0f;
            qw = (r02 - r20) / s4; qx = (r01 + r10) / s4; qy = 0.25f * s4; qz = (r12 + r21) / s4;
        } else {
            const float s4 = sqrtf(1.0f + r22 - r00 - r11) * 2.0f;
            qw = (r10 - r01) / s4; qx = (r02 + r20) / s4; qy = (r12 + r21) / s4; qz = 0.25f * s4;
        }
        const float inv = 1.0f / sqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
        qx *= inv; qy *= inv; qz *= inv; qw *= inv;
        const float tx = col[0].w, ty = col[1].w, tz = col[2].w;
        const float dx = 0.5f * (tx * qw + ty * qz - tz * qy);
        const float dy = 0.5f * (-tx * qz + ty * qw + tz * qx);
        const float dz = 0.5f * (tx * qy - ty * qx + tz * qw);
        const float dw = -0.5f * (tx * qx + ty * qy + tz * qz);
        S.pal_ext[2 * (size_t)b + 0] = make_float4(qx, qy, qz, qw);
        S.pal_ext[2 * (size_t)b + 1] = make_float4(dx, dy, dz, dw);
    }
}

// Extension (parity unpinned): material morph accumulation.  libmmd allocates Poser::material_mul_images_ (all 1)
// and material_add_images_ (all 0) (L/motion/poser.inl:160-161, poser_impl.inl:31-36) but UpdateMorphTransform
// never fills them (poser_impl.inl:355-358).  Here, in application order per material and per field:
//   MUL entry:  mul = mul * (1 + (value - 1) * rate)       ADD entry:  add = add + value * rate
// (a renderer shows base * mul + add).  Skipped morphs have rate 0 and leave both images unchanged.
__device__ __forceinline__ void accumulate_material_images(const DevModel& M, const DevFrames& F, uint32_t slot,
                                                           const float* nrate, uint32_t tid, uint32_t nthreads) {
    if (!F.material_images) return;
    constexpr uint32_t NF = MMDGPU_MATERIAL_FIELDS;
    float* out = F.material_images + (size_t)slot * M.n_materials * 2 * NF;
    for (uint32_t i = tid; i < M.n_materials * NF; i += nthreads) {
        const uint32_t mat = i / NF, k = i % NF;
        float mul = 1.0f, add = 0.0f;
        for (int32_t e = M.material_morph_row[mat]; e < M.material_morph_row[mat + 1]; ++e) {
            const MaterialMorphEntry* E = M.material_morph_entries + e;
            const float r = nrate[kSlotGroup * E->node];
            if (r != 0.0f) {
                const float v = __ldg(&E->value[k]);
                if (E->method == MMDGPU_MATERIAL_MUL) mul = mul * (1.0f + (v - 1.0f) * r);
                else add = add + v * r;
            }
        }
        out[(size_t)mat * 2 * NF + k] = mul;
        out[(size_t)mat * 2 * NF + NF + k] = add;
    }
}

constexpr uint32_t kHierWarps = 4;

// one slot's bone state where the segment kernels leave it between launches: global memory
__device__ __forceinline__ SlotState global_slot_state(const DevModel& M, const DevFrames& F, uint32_t slot) {
    SlotState S;
    S.bones = M.bones;
    S.poseR = F.poseR + (size_t)slot * M.nb;
    S.poseT = F.poseT + (size_t)slot * M.nb;
    S.totR = F.totR + (size_t)slot * M.nb;
    S.totT = F.totT + (size_t)slot * M.nb;
    S.local = reinterpret_cast<float4*>(F.local) + (size_t)slot * M.nb * 3;
    S.ikR = F.ikR + (size_t)slot * M.n_link_slots;
    S.preIK = F.preIK + (size_t)slot * M.n_link_slots;
    S.morphR = F.morphR + (size_t)slot * M.n_morph_slots;
    S.morphT = F.morphT + (size_t)slot * M.n_morph_slots;
    S.palette = F.palette + (size_t)slot * M.nb * 3;
    S.pal_ext = F.pal_ext ? F.pal_ext + (size_t)slot * M.nb * 2 : nullptr;
    return S;
}

// -------------------------------------------------------------------------------------------------
// K2, flat form of ONE wave: a thread per (op of the wave, slot), slot-minor.  Used for the waves that hold CCD IK
// solves when many slots are evaluated together.  In the CTA-per-slot kernel a solve keeps a CTA (34 KB of shared
// memory on a 200-bone rig) resident for 40 x links dependent steps of one or two active lanes; here 32 slots'
// solves of the same chain share a warp and each solve works on a CHAIN-LOCAL IMAGE of the state (IkImage: links,
// target, their parents, < 2 KB) in shared memory, copied in from and back to the global state the segment kernels
// exchange.  The image carries translated copies of the static records, so eval_bone / solve_ik run unchanged:
// same arithmetic, bit-identical results.  Non-IK ops that share the wave run on the global state.
// -------------------------------------------------------------------------------------------------
constexpr uint32_t kFlatThreads = 64;

__global__ void __launch_bounds__(kFlatThreads) hierarchy_flat_kernel(DevModel M, DevFrames F, uint32_t wave) {
    extern __shared__ __align__(16) float4 fsm[];
    const uint32_t o0 = M.wave_begin[wave], n = M.wave_begin[wave + 1] - o0;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * F.n_slots) return;
    const uint32_t slot = t % F.n_slots;
    const SlotState G = global_slot_state(M, F, slot);
    const uint32_t word = __ldg(M.wave_ops + o0 + t / F.n_slots);
    const uint32_t kind = word >> 28;
    const int32_t arg = (int32_t)(word & 0x0FFFFFFFu);
    if (kind == kOpEval) { eval_bone(M, G, arg); return; }
    if (kind == kOpSkin) { skin_bone(M, G, arg); return; }
    const IkImage I = M.ik_img[arg];
    float4* reg = fsm + (size_t)threadIdx.x * M.ik_img_max_region;
    __builtin_assume(__isShared(reg));
    const uint32_t W = (uint32_t)I.n_bones, L = (uint32_t)I.n_lslots, Ms = (uint32_t)I.n_mslots;
    SlotState S;
    S.bones = M.ik_img_static + I.bones_begin;
    float4* totR = reg;
    float4* totT = totR + W;
    float4* local = totT + W;            // 3 per bone
    float4* poseR = local + 3 * W;
    float4* poseT = poseR + W;
    float4* ikR = poseT + W;
    float4* preIK = ikR + L;
    float4* morphR = preIK + L;
    float4* morphT = morphR + Ms;
    S.totR = totR; S.totT = totT; S.local = local; S.poseR = poseR; S.poseT = poseT; S.ikR = ikR; S.preIK = preIK;
    S.morphR = morphR; S.morphT = morphT;
    S.palette = nullptr; S.pal_ext = nullptr;
    const int32_t* __restrict__ gb = M.ik_img_bones + I.bones_begin;
    for (uint32_t i = 0; i < W; ++i) {
        const int32_t b = __ldg(gb + i);
        totR[i] = G.totR[b];
        totT[i] = G.totT[b];
        local[3 * i] = G.local[3 * (size_t)b];
        local[3 * i + 1] = G.local[3 * (size_t)b + 1];
        local[3 * i + 2] = G.local[3 * (size_t)b + 2];
        poseR[i] = G.poseR[b];
        poseT[i] = G.poseT[b];
    }
    const int32_t* __restrict__ gl = M.ik_img_lslots + I.lslots_begin;
    for (uint32_t i = 0; i < L; ++i) { const int32_t x = __ldg(gl + i); ikR[i] = G.ikR[x]; preIK[i] = G.preIK[x]; }
    const int32_t* __restrict__ gm = M.ik_img_mslots + I.mslots_begin;
    for (uint32_t i = 0; i < Ms; ++i) { const int32_t x = __ldg(gm + i); morphR[i] = G.morphR[x]; morphT[i] = G.morphT[x]; }

    solve_ik<0, false>(IkTables{M.ik_img_desc, M.ik_img_links}, S, M.ik_img_desc[arg]);

    const uint8_t* __restrict__ wr = M.ik_img_written + I.bones_begin;
    for (uint32_t i = 0; i < W; ++i) {
        if (!__ldg(wr + i)) continue;
        const int32_t b = __ldg(gb + i);
        G.totR[b] = totR[i];
        G.totT[b] = totT[i];
        G.local[3 * (size_t)b] = local[3 * i];
        G.local[3 * (size_t)b + 1] = local[3 * i + 1];
        G.local[3 * (size_t)b + 2] = local[3 * i + 2];
    }
    for (uint32_t i = 0; i < L; ++i) { const int32_t x = __ldg(gl + i); G.ikR[x] = ikR[i]; G.preIK[x] = preIK[i]; }
}

template <bool NEST>
__global__ void __launch_bounds__(32 * kHierWarps) hierarchy_kernel(DevModel M, DevFrames F, uint32_t wave_lo,
                                                                    uint32_t wave_hi, uint32_t prologue) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t slot = blockIdx.x * kHierWarps + (threadIdx.x >> 5);
    if (slot >= F.n_slots) return;
    const SlotState S = global_slot_state(M, F, slot);
    float4* morphR = F.morphR + (size_t)slot * M.n_morph_slots;
    float4* morphT = F.morphT + (size_t)slot * M.n_morph_slots;

    if (prologue) {
        // ---- morph application-slot rates: Poser::UpdateMorphTransform's skip test and group recursion
        //      (poser_impl.inl:329-339), evaluated breadth-first over the static DFS tree.
        const float* rate = F.rate + (size_t)slot * M.nm;
        // application-slot rates are stored [slot / 4][node][slot % 4]: the skinning kernel evaluates four slots
        // together and reads one float4 per morph entry
        float* nrate = F.node_rate + (size_t)(slot >> 2) * M.n_nodes_pad * kSlotGroup + (slot & 3u);
        for (uint32_t dpt = 0; dpt < M.n_depths; ++dpt) {
            const int32_t b0 = M.depth_begin[dpt], b1 = M.depth_begin[dpt + 1];
            for (int32_t i = b0 + (int32_t)lane; i < b1; i += 32) {
                const int32_t n = M.nodes_by_depth[i];
                const int32_t par = M.node_parent[n];
                float r;
                bool active = true;
                if (par < 0) r = rate[M.node_morph[n]];
                else {
                    const float pr = nrate[kSlotGroup * par];
                    active = pr != 0.0f;            // a skipped group skips its whole subtree
                    r = M.node_mult[n] * pr;        // data.GetMorphRate()*rate
                }
                if (!active || (double)r < kEpsD) r = 0.0f;
                nrate[kSlotGroup * n] = r;
            }
            __syncwarp();
        }
        // ---- PrePhysicsPosing's per-bone reset (poser_impl.inl:366-377), restricted to state that is read
        //      before it is written this frame, and the compact IK / bone-morph state.
        for (uint32_t i = lane; i < M.n_reset; i += 32) {
            const int32_t b = M.reset_bones[i];
            S.totR[b] = make_float4(0.f, 0.f, 0.f, 1.f);
            S.totT[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            store_local(S.local + 3 * (size_t)b, m_identity());
        }
        for (uint32_t i = lane; i < M.n_link_slots; i += 32) {
            S.ikR[i] = make_float4(0.f, 0.f, 0.f, 1.f);
            S.preIK[i] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
        // ---- bone morphs (poser_impl.inl:347-354), application order inside each affected bone
        for (uint32_t i = lane; i < M.n_morph_slots; i += 32) {
            Quat mr = q_identity();
            float tx = 0.f, ty = 0.f, tz = 0.f;
            for (int32_t e = M.bone_morph_row[i]; e < M.bone_morph_row[i + 1]; ++e) {
                const BoneMorphEntry E = M.bone_morph_entries[e];
                const float r = nrate[kSlotGroup * E.node];
                if (r != 0.0f) {
                    tx = tx + E.translation[0] * r;
                    ty = ty + E.translation[1] * r;
                    tz = tz + E.translation[2] * r;
                    const Quat q{E.rotation[0], E.rotation[1], E.rotation[2], E.rotation[3]};
                    mr = q_mul(mr, q_slerp(q_identity(), q, r));
                }
            }
            morphR[i] = q_to4(mr);
            morphT[i] = make_float4(tx, ty, tz, 0.f);
        }
        accumulate_material_images(M, F, slot, nrate, lane, 32);
        __syncwarp();
    }

    for (uint32_t w = wave_lo; w < wave_hi; ++w) {
        const uint32_t o0 = M.wave_begin[w], o1 = M.wave_begin[w + 1];
        for (uint32_t o = o0 + lane; o < o1; o += 32) {
            const uint32_t word = __ldg(M.wave_ops + o);
            const uint32_t kind = word >> 28;
            const int32_t arg = (int32_t)(word & 0x0FFFFFFFu);
            if (kind == kOpEval) eval_bone(M, S, arg);
            else if (kind == kOpIk) solve_ik<0, NEST>(IkTables{M.iks, M.links}, S, M.iks[arg]);
            else skin_bone(M, S, arg);
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// K2, shared-memory form: one CTA per slot keeps the slot's whole bone state (poses, total rotation /
// translation, local matrices, IK and bone-morph scratch: 112 B per bone) in shared memory, so the dependent chain
// of the wave program runs at shared-memory latency instead of a global-memory round trip per wave.  Used whenever
// the state fits (about 2000 bones); the warp-per-slot kernel above remains the fallback.
// -------------------------------------------------------------------------------------------------
constexpr uint32_t kHierCtaThreads = 256;
constexpr size_t kHierCtaSmemLimit = 200 * 1024;

// per bone: poses 2 + totals 2 + local 3 + static record 3 float4; plus the op words of the program
__host__ __device__ inline size_t hier_cta_smem_bytes(uint32_t nb, uint32_t n_link_slots, uint32_t n_morph_slots, uint32_t n_ops,
                                                      uint32_t n_waves) {
    return ((size_t)nb * 10 + (size_t)n_link_slots * 2 + (size_t)n_morph_slots * 2) * sizeof(float4) +
           (((size_t)n_ops + n_waves + 1 + 3) & ~(size_t)3) * sizeof(uint32_t);
}

template <bool NEST>
__global__ void __launch_bounds__(kHierCtaThreads) hierarchy_cta_kernel(DevModel M, DevFrames F, uint32_t wave_lo,
                                                                        uint32_t wave_hi, uint32_t prologue) {
    extern __shared__ __align__(16) float4 hsm[];
    const uint32_t slot = blockIdx.x, tid = threadIdx.x, nb = M.nb, nthreads = blockDim.x;
    float4* s_poseR = hsm;
    float4* s_poseT = s_poseR + nb;
    float4* s_totR = s_poseT + nb;
    float4* s_totT = s_totR + nb;
    float4* s_local = s_totT + nb;  // 3 per bone
    float4* s_ikR = s_local + 3 * (size_t)nb;
    float4* s_preIK = s_ikR + M.n_link_slots;
    float4* s_morphR = s_preIK + M.n_link_slots;
    float4* s_morphT = s_morphR + M.n_morph_slots;
    float4* s_bones = s_morphT + M.n_morph_slots;  // 3 per bone: the static records
    uint32_t* s_wave_begin = reinterpret_cast<uint32_t*>(s_bones + 3 * (size_t)nb);
    uint32_t* s_wave_ops = s_wave_begin + M.n_waves + 1;

    float4* g_totR = F.totR + (size_t)slot * nb;
    float4* g_totT = F.totT + (size_t)slot * nb;
    float4* g_local = reinterpret_cast<float4*>(F.local) + (size_t)slot * nb * 3;
    float4* g_ikR = F.ikR + (size_t)slot * M.n_link_slots;
    float4* g_preIK = F.preIK + (size_t)slot * M.n_link_slots;
    float4* g_morphR = F.morphR + (size_t)slot * M.n_morph_slots;
    float4* g_morphT = F.morphT + (size_t)slot * M.n_morph_slots;

    // tell the compiler these pointers are shared-memory addresses: plain LDS / STS instead of generic accesses
    __builtin_assume(__isShared(s_poseR)); __builtin_assume(__isShared(s_poseT)); __builtin_assume(__isShared(s_totR));
    __builtin_assume(__isShared(s_totT)); __builtin_assume(__isShared(s_local)); __builtin_assume(__isShared(s_ikR));
    __builtin_assume(__isShared(s_preIK)); __builtin_assume(__isShared(s_morphR)); __builtin_assume(__isShared(s_morphT));
    __builtin_assume(__isShared(s_bones));
    SlotState S;
    S.bones = reinterpret_cast<const BoneStatic*>(s_bones);
    S.poseR = s_poseR; S.poseT = s_poseT; S.totR = s_totR; S.totT = s_totT; S.local = s_local;
    S.ikR = s_ikR; S.preIK = s_preIK; S.morphR = s_morphR; S.morphT = s_morphT;
    S.palette = F.palette + (size_t)slot * nb * 3;
    S.pal_ext = F.pal_ext ? F.pal_ext + (size_t)slot * nb * 2 : nullptr;

    // ---- the static bone records and the program (model data: fetched while a predecessor launched with programmatic
    //      dependent launch may still be running), then the sampled poses of every bone (written by K1 / SetBonePose)
    pdl_trigger();
    {
        const float4* gb = reinterpret_cast<const float4*>(M.bones);
        for (uint32_t i = tid; i < 3 * nb; i += nthreads) s_bones[i] = __ldg(gb + i);
        for (uint32_t i = tid; i <= M.n_waves; i += nthreads) s_wave_begin[i] = __ldg(M.wave_begin + i);
        const uint32_t n_ops = __ldg(M.wave_begin + M.n_waves);
        for (uint32_t i = tid; i < n_ops; i += nthreads) s_wave_ops[i] = __ldg(M.wave_ops + i);
    }
    pdl_wait();
    for (uint32_t b = tid; b < nb; b += nthreads) {
        s_poseR[b] = F.poseR[(size_t)slot * nb + b];
        s_poseT[b] = F.poseT[(size_t)slot * nb + b];
    }
    if (prologue) {
        // ---- morph application-slot rates (poser_impl.inl:329-339), breadth-first over the static DFS tree
        const float* rate = F.rate + (size_t)slot * M.nm;
        // application-slot rates are stored [slot / 4][node][slot % 4]: the skinning kernel evaluates four slots
        // together and reads one float4 per morph entry
        float* nrate = F.node_rate + (size_t)(slot >> 2) * M.n_nodes_pad * kSlotGroup + (slot & 3u);
        for (uint32_t dpt = 0; dpt < M.n_depths; ++dpt) {
            const int32_t b0 = M.depth_begin[dpt], b1 = M.depth_begin[dpt + 1];
            for (int32_t i = b0 + (int32_t)tid; i < b1; i += nthreads) {
                const int32_t n = M.nodes_by_depth[i];
                const int32_t par = M.node_parent[n];
                float r;
                bool active = true;
                if (par < 0) r = rate[M.node_morph[n]];
                else {
                    const float pr = nrate[kSlotGroup * par];
                    active = pr != 0.0f;
                    r = M.node_mult[n] * pr;
                }
                if (!active || (double)r < kEpsD) r = 0.0f;
                nrate[kSlotGroup * n] = r;
            }
            if (dpt + 1 < M.n_depths) { __threadfence_block(); __syncthreads(); }
        }
        // ---- PrePhysicsPosing's per-bone reset (poser_impl.inl:366-377)
        for (uint32_t b = tid; b < nb; b += nthreads) {
            s_totR[b] = make_float4(0.f, 0.f, 0.f, 1.f);
            s_totT[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            store_local(s_local + 3 * (size_t)b, m_identity());
        }
        for (uint32_t i = tid; i < M.n_link_slots; i += nthreads) {
            s_ikR[i] = make_float4(0.f, 0.f, 0.f, 1.f);
            s_preIK[i] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
        __syncthreads();  // node rates of this CTA are visible (written by its own threads)
        // ---- bone morphs (poser_impl.inl:347-354), application order inside each affected bone
        for (uint32_t i = tid; i < M.n_morph_slots; i += nthreads) {
            Quat mr = q_identity();
            float tx = 0.f, ty = 0.f, tz = 0.f;
            for (int32_t e = M.bone_morph_row[i]; e < M.bone_morph_row[i + 1]; ++e) {
                const BoneMorphEntry E = M.bone_morph_entries[e];
                const float r = nrate[kSlotGroup * E.node];
                if (r != 0.0f) {
                    tx = tx + E.translation[0] * r;
                    ty = ty + E.translation[1] * r;
                    tz = tz + E.translation[2] * r;
                    const Quat q{E.rotation[0], E.rotation[1], E.rotation[2], E.rotation[3]};
                    mr = q_mul(mr, q_slerp(q_identity(), q, r));
                }
            }
            s_morphR[i] = q_to4(mr);
            s_morphT[i] = make_float4(tx, ty, tz, 0.f);
        }
        accumulate_material_images(M, F, slot, nrate, tid, nthreads);
    } else {
        // continue from the state the previous launch (pre-physics segment, possibly edited by the host physics
        // hook) left in global memory
        for (uint32_t b = tid; b < nb; b += nthreads) {
            s_totR[b] = g_totR[b];
            s_totT[b] = g_totT[b];
            s_local[3 * (size_t)b] = g_local[3 * (size_t)b];
            s_local[3 * (size_t)b + 1] = g_local[3 * (size_t)b + 1];
            s_local[3 * (size_t)b + 2] = g_local[3 * (size_t)b + 2];
        }
        for (uint32_t i = tid; i < M.n_link_slots; i += nthreads) { s_ikR[i] = g_ikR[i]; s_preIK[i] = g_preIK[i]; }
        for (uint32_t i = tid; i < M.n_morph_slots; i += nthreads) { s_morphR[i] = g_morphR[i]; s_morphT[i] = g_morphT[i]; }
    }
    __syncthreads();

    for (uint32_t w = wave_lo; w < wave_hi; ++w) {
        const uint32_t o0 = s_wave_begin[w], o1 = s_wave_begin[w + 1];
        for (uint32_t o = o0 + tid; o < o1; o += nthreads) {
            const uint32_t word = s_wave_ops[o];
            const uint32_t kind = word >> 28;
            const int32_t arg = (int32_t)(word & 0x0FFFFFFFu);
            if (kind == kOpEval) eval_bone(M, S, arg);
            else if (kind == kOpIk) solve_ik<0, NEST>(IkTables{M.iks, M.links}, S, M.iks[arg]);
            else skin_bone(M, S, arg);
        }
        __syncthreads();
    }

    // ---- leave the state in global memory for the next segment / the download entry points
    for (uint32_t b = tid; b < nb; b += nthreads) {
        g_totR[b] = s_totR[b];
        g_totT[b] = s_totT[b];
        g_local[3 * (size_t)b] = s_local[3 * (size_t)b];
        g_local[3 * (size_t)b + 1] = s_local[3 * (size_t)b + 1];
        g_local[3 * (size_t)b + 2] = s_local[3 * (size_t)b + 2];
    }
    for (uint32_t i = tid; i < M.n_link_slots; i += nthreads) { g_ikR[i] = s_ikR[i]; g_preIK[i] = s_preIK[i]; }
    for (uint32_t i = tid; i < M.n_morph_slots; i += nthreads) { g_morphR[i] = s_morphR[i]; g_morphT[i] = s_morphT[i]; }
}

// =================================================================================================
// K3 — skinning.  Poser::Deform, L/motion/poser_impl.inl:396-461; transform / rotate math_impl.inl:1032-1045.
//
// Work item = one 512-vertex tile x a run of consecutive slots (frames of a bake, instances of a crowd), walked
// four slots (one "slot group") at a time.
//   * The tile's static streams are read ONCE (16-byte coalesced loads, 4 storage positions per thread) and stay
//     in registers while the CTA walks its slots: per vertex-frame only the output leaves the SM.
//   * Per slot group the CTA stages just the matrices of the bones this tile uses (tile-local palettes of the four
//     slots) and the group's morph application-slot rates (one float4 per slot = the four slots' rates) in shared
//     memory, double-buffered: the next group's loads are in flight while the current group is computed.
//   * Sparse vertex morphs arrive as a sliced-ELL gather (32-lane groups, coalesced 512-byte rounds, warp-uniform
//     trip count) in libmmd's application order; one entry load serves the four slots.  No vertex_images_ buffer,
//     no clear pass, no atomics.
//   * Tiles are stored sorted by (skinning type, morph entry count), so a warp step runs one branch; results are
//     written to shared-memory staging tiles at the vertex's PMX index and leave the SM as one bulk async copy
//     (cp.async.bulk shared -> global) per output plane and slot.  The interleaved layout's 32-byte records are
//     whole DRAM sectors and go straight to global memory, one 256-bit store each.
//   * The blend + transform is a real function (skin_vertex_n: two slots of one vertex per call); inlining every
//     call site thrashed the instruction cache.
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct Col3 { float4 c0, c1, c2; };  // skinning matrix as three columns (M0c, M1c, M2c, M3c)

template <int PS>  // PS = float4 per staged bone: 3 (matrix columns) or 5 (+ rotation quaternion, dual part)
__device__ __forceinline__ Col3 pal_load(const float4* __restrict__ pal, uint32_t id) {
    Col3 r;
    r.c0 = pal[PS * id + 0];
    r.c1 = pal[PS * id + 1];
    r.c2 = pal[PS * id + 2];
    return r;
}
__device__ __forceinline__ float4 f4_scale(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// One vertex.  ids: 4 x u16 tile-local bone indices (type in bits 15:13 of id0); w: BDEF2 uses w.x, BDEF4 all four.
// Deliberately NOT inlined: the skinning kernel calls it V x G = 16 times per slot group, and the fully inlined
// kernel stalled a quarter of its issue slots on instruction fetch.  Results come back in registers.
// One 32-byte sokol vertex (main.cpp:50-54) as a single 256-bit streaming store (sm_100: STG.E.256): the record is
// exactly one DRAM sector, written by one request instead of two 16-byte halves.
__device__ __forceinline__ void store_record32(float4* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e),
                 "f"(f), "f"(g), "f"(h)
                 : "memory");
}

struct Skinned { float px, py, pz, nx, ny, nz; };
template <int PS, bool PAL_SHARED>
__device__ __noinline__ Skinned skin_vertex(const float4* __restrict__ pal, uint32_t ids_lo, uint32_t ids_hi, float4 w,
                                            float px, float py, float pz, float nx, float ny, float nz) {
    if (PAL_SHARED) __builtin_assume(__isShared(pal));  // staged palette: shared-memory loads, not generic ones
    const uint32_t type = (ids_lo >> 13) & 7u;
    const uint32_t id0 = ids_lo & 0x1FFFu, id1 = ids_lo >> 16;
    Col3 Mx = pal_load<PS>(pal, id0);
    if (type == kDevBdef2) {
        // Lerp(mat_1, mat_0)[w] = (1-w)*mat_1 + w*mat_0 (poser_impl.inl:422, math_impl.inl:1246-1254)
        const Col3 B = pal_load<PS>(pal, id1);
        const float l = w.x, om = 1.0f - w.x;
        Mx.c0 = f4_add(f4_scale(B.c0, om), f4_scale(Mx.c0, l));
        Mx.c1 = f4_add(f4_scale(B.c1, om), f4_scale(Mx.c1, l));
        Mx.c2 = f4_add(f4_scale(B.c2, om), f4_scale(Mx.c2, l));
    } else if (type == kDevBdef4) {
        // mat_0*w0 + mat_1*w1 + mat_2*w2 + mat_3*w3, left to right (poser_impl.inl:433)
        const uint32_t id2 = ids_hi & 0xFFFFu, id3 = ids_hi >> 16;
        const Col3 B = pal_load<PS>(pal, id1), C = pal_load<PS>(pal, id2), D = pal_load<PS>(pal, id3);
        Mx.c0 = f4_add(f4_add(f4_add(f4_scale(Mx.c0, w.x), f4_scale(B.c0, w.y)), f4_scale(C.c0, w.z)), f4_scale(D.c0, w.w));
        Mx.c1 = f4_add(f4_add(f4_add(f4_scale(Mx.c1, w.x), f4_scale(B.c1, w.y)), f4_scale(C.c1, w.z)), f4_scale(D.c1, w.w));
        Mx.c2 = f4_add(f4_add(f4_add(f4_scale(Mx.c2, w.x), f4_scale(B.c2, w.y)), f4_scale(C.c2, w.z)), f4_scale(D.c2, w.w));
    }
    // transform: v0*m00 + v1*m10 + v2*m20 + m30 ; rotate: without the translation row
    Skinned r;
    r.px = px * Mx.c0.x + py * Mx.c0.y + pz * Mx.c0.z + Mx.c0.w;
    r.py = px * Mx.c1.x + py * Mx.c1.y + pz * Mx.c1.z + Mx.c1.w;
    r.pz = px * Mx.c2.x + py * Mx.c2.y + pz * Mx.c2.z + Mx.c2.w;
    r.nx = nx * Mx.c0.x + ny * Mx.c0.y + nz * Mx.c0.z;
    r.ny = nx * Mx.c1.x + ny * Mx.c1.y + nz * Mx.c1.z;
    r.nz = nx * Mx.c2.x + ny * Mx.c2.y + nz * Mx.c2.z;
    return r;
}

// NS slots of one vertex in one call: the id / weight decode is shared and the NS blends are independent
// instruction streams the scheduler can interleave.  pal points at slot 0's palette, the others follow at pal4.
template <int NS> struct SkinnedN { Skinned s[NS]; };
template <int NS> struct MorphedN { float x[NS], y[NS], z[NS]; };
template <int PS, bool PAL_SHARED, int NS>
__device__ __noinline__ SkinnedN<NS> skin_vertex_n(const float4* __restrict__ pal, uint32_t pal4, uint32_t ids_lo, uint32_t ids_hi,
                                                   float4 w, MorphedN<NS> q, float nx, float ny, float nz) {
    if (PAL_SHARED) __builtin_assume(__isShared(pal));
    const uint32_t type = (ids_lo >> 13) & 7u;
    const uint32_t id0 = ids_lo & 0x1FFFu, id1 = ids_lo >> 16;
    Col3 Mx[NS];
#pragma unroll
    for (int f = 0; f < NS; ++f) Mx[f] = pal_load<PS>(pal + (size_t)f * pal4, id0);
    if (type == kDevBdef2) {
        const float l = w.x, om = 1.0f - w.x;
#pragma unroll
        for (int f = 0; f < NS; ++f) {
            const Col3 B = pal_load<PS>(pal + (size_t)f * pal4, id1);
            Mx[f].c0 = f4_add(f4_scale(B.c0, om), f4_scale(Mx[f].c0, l));
            Mx[f].c1 = f4_add(f4_scale(B.c1, om), f4_scale(Mx[f].c1, l));
            Mx[f].c2 = f4_add(f4_scale(B.c2, om), f4_scale(Mx[f].c2, l));
        }
    } else if (type == kDevBdef4) {
        const uint32_t id2 = ids_hi & 0xFFFFu, id3 = ids_hi >> 16;
#pragma unroll
        for (int f = 0; f < NS; ++f) {
            const float4* pf = pal + (size_t)f * pal4;
            const Col3 B = pal_load<PS>(pf, id1), C = pal_load<PS>(pf, id2), D = pal_load<PS>(pf, id3);
            Mx[f].c0 = f4_add(f4_add(f4_add(f4_scale(Mx[f].c0, w.x), f4_scale(B.c0, w.y)), f4_scale(C.c0, w.z)), f4_scale(D.c0, w.w));
            Mx[f].c1 = f4_add(f4_add(f4_add(f4_scale(Mx[f].c1, w.x), f4_scale(B.c1, w.y)), f4_scale(C.c1, w.z)), f4_scale(D.c1, w.w));
            Mx[f].c2 = f4_add(f4_add(f4_add(f4_scale(Mx[f].c2, w.x), f4_scale(B.c2, w.y)), f4_scale(C.c2, w.z)), f4_scale(D.c2, w.w));
        }
    }
    SkinnedN<NS> r;
#pragma unroll
    for (int f = 0; f < NS; ++f) {
        const Col3& A = Mx[f];
        r.s[f].px = q.x[f] * A.c0.x + q.y[f] * A.c0.y + q.z[f] * A.c0.z + A.c0.w;
        r.s[f].py = q.x[f] * A.c1.x + q.y[f] * A.c1.y + q.z[f] * A.c1.z + A.c1.w;
        r.s[f].pz = q.x[f] * A.c2.x + q.y[f] * A.c2.y + q.z[f] * A.c2.z + A.c2.w;
        r.s[f].nx = nx * A.c0.x + ny * A.c0.y + nz * A.c0.z;
        r.s[f].ny = nx * A.c1.x + ny * A.c1.y + nz * A.c1.z;
        r.s[f].nz = nx * A.c2.x + ny * A.c2.y + nz * A.c2.z;
    }
    return r;
}

// ---- extensions (parity unpinned: libmmd implements none of these) -----------------------------
__device__ __forceinline__ void quat_rotate(const float4& q, float vx, float vy, float vz, float* o) {
    // v' = v + 2 * cross(q.xyz, cross(q.xyz, v) + q.w * v)
    const float cx = q.y * vz - q.z * vy + q.w * vx, cy = q.z * vx - q.x * vz + q.w * vy, cz = q.x * vy - q.y * vx + q.w * vz;
    o[0] = vx + 2.0f * (q.y * cz - q.z * cy);
    o[1] = vy + 2.0f * (q.z * cx - q.x * cz);
    o[2] = vz + 2.0f * (q.x * cy - q.y * cx);
}
// Spherical deform (PMX SDEF): rotate about the centre C with slerp(q0, q1, w1), translate by the weighted
// transformed centres cr0 / cr1 (precomputed at load).  sd = {C, cr0, cr1}.
__device__ __forceinline__ void skin_sdef(const float4* __restrict__ pal, uint32_t id0, uint32_t id1, float w0,
                                          const float4* __restrict__ sd, float px, float py, float pz, float nx, float ny,
                                          float nz, float* __restrict__ op, float* __restrict__ on) {
    const float w1 = 1.0f - w0;
    const Col3 A = pal_load<5>(pal, id0), B = pal_load<5>(pal, id1);
    const float4 q0 = pal[5 * id0 + 3];
    float4 q1 = pal[5 * id1 + 3];
    float dot = q0.x * q1.x + q0.y * q1.y + q0.z * q1.z + q0.w * q1.w;
    if (dot < 0.0f) { q1 = make_float4(-q1.x, -q1.y, -q1.z, -q1.w); dot = -dot; }
    float k0 = w0, k1 = w1;
    if (dot < 0.9995f) {
        const float om = acosf(dot), so = sinf(om);
        k0 = sinf(w0 * om) / so;
        k1 = sinf(w1 * om) / so;
    }
    float4 q = make_float4(q0.x * k0 + q1.x * k1, q0.y * k0 + q1.y * k1, q0.z * k0 + q1.z * k1, q0.w * k0 + q1.w * k1);
    const float inv = 1.0f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
    const float4 C = sd[0], c0 = sd[1], c1 = sd[2];
    float r[3];
    quat_rotate(q, px - C.x, py - C.y, pz - C.z, r);
    const float t0x = c0.x * A.c0.x + c0.y * A.c0.y + c0.z * A.c0.z + A.c0.w, t0y = c0.x * A.c1.x + c0.y * A.c1.y + c0.z * A.c1.z + A.c1.w,
                t0z = c0.x * A.c2.x + c0.y * A.c2.y + c0.z * A.c2.z + A.c2.w;
    const float t1x = c1.x * B.c0.x + c1.y * B.c0.y + c1.z * B.c0.z + B.c0.w, t1y = c1.x * B.c1.x + c1.y * B.c1.y + c1.z * B.c1.z + B.c1.w,
                t1z = c1.x * B.c2.x + c1.y * B.c2.y + c1.z * B.c2.z + B.c2.w;
    op[0] = r[0] + t0x * w0 + t1x * w1;
    op[1] = r[1] + t0y * w0 + t1y * w1;
    op[2] = r[2] + t0z * w0 + t1z * w1;
    quat_rotate(q, nx, ny, nz, on);
}
// Dual-quaternion blend (PMX 2.1 QDEF) of four bones, antipodality resolved against the first bone.
__device__ __forceinline__ void skin_qdef(const float4* __restrict__ pal, uint32_t ids_lo, uint32_t ids_hi, const float4& w,
                                          float px, float py, float pz, float nx, float ny, float nz, float* __restrict__ op,
                                          float* __restrict__ on) {
    const uint32_t id[4] = {ids_lo & 0x1FFFu, ids_lo >> 16, ids_hi & 0xFFFFu, ids_hi >> 16};
    const float wt[4] = {w.x, w.y, w.z, w.w};
    const float4 qa = pal[5 * id[0] + 3];
    float4 br = make_float4(0.f, 0.f, 0.f, 0.f), bd = br;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 q = pal[5 * id[i] + 3], d = pal[5 * id[i] + 4];
        float s = wt[i];
        if (q.x * qa.x + q.y * qa.y + q.z * qa.z + q.w * qa.w < 0.0f) s = -s;
        br = make_float4(br.x + q.x * s, br.y + q.y * s, br.z + q.z * s, br.w + q.w * s);
        bd = make_float4(bd.x + d.x * s, bd.y + d.y * s, bd.z + d.z * s, bd.w + d.w * s);
    }
    const float inv = 1.0f / sqrtf(br.x * br.x + br.y * br.y + br.z * br.z + br.w * br.w);
    br = make_float4(br.x * inv, br.y * inv, br.z * inv, br.w * inv);
    bd = make_float4(bd.x * inv, bd.y * inv, bd.z * inv, bd.w * inv);
    float r[3];
    quat_rotate(br, px, py, pz, r);
    // translation of a unit dual quaternion: 2 * (qr.w * qd.xyz - qd.w * qr.xyz + cross(qr.xyz, qd.xyz))
    op[0] = r[0] + 2.0f * (br.w * bd.x - bd.w * br.x + (br.y * bd.z - br.z * bd.y));
    op[1] = r[1] + 2.0f * (br.w * bd.y - bd.w * br.y + (br.z * bd.x - br.x * bd.z));
    op[2] = r[2] + 2.0f * (br.w * bd.z - bd.w * br.z + (br.x * bd.y - br.y * bd.x));
    quat_rotate(br, nx, ny, nz, on);
}

// shared memory carve-up (bytes): [stage 0][stage 1][palette 0][palette 1][rates 0][rates 1]
// Staging tile per slot.  The sokol32 layout needs none: its 32-byte records are whole DRAM sectors, so threads store
// them straight to global memory (every sector is written exactly once, in full); the 12-byte SoA records would be
// partial-sector writes and go through a shared-memory tile + bulk copy instead.
__host__ __device__ inline uint32_t skin_stage_bytes(int layout, bool ext) {
    return layout == MMDGPU_LAYOUT_SOA_POS_NRM ? kTileVerts * (ext ? 32u : 24u) : 0u;  // ext SoA: + UV plane
}
// the packed-pair kernel also stages the 32-byte records (16 KB per slot) unless told not to (experiment knob)
__host__ __device__ constexpr uint32_t skin_pair_stage_bytes(int layout, bool sokol_staged) {
    return layout == MMDGPU_LAYOUT_SOA_POS_NRM ? kTileVerts * 24u : (sokol_staged ? kTileVerts * 32u : 0u);
}
__host__ __device__ inline uint32_t skin_pal_bytes(uint32_t max_tile_bones, bool ext) { return max_tile_bones * (ext ? 80u : 48u); }

constexpr uint32_t kPalPrefetch = 2;  // palette float4 per thread held in registers across the compute phase
constexpr int V = (int)kVertsPerThread;
constexpr int G = (int)kSlotGroup;    // slots one CTA evaluates together

// shared memory carve-up (bytes): [stage: G tiles][palette 0: G slots][palette 1][rates 0: n_nodes_pad float4][rates 1]
// PALG: global-palette models (a tile touches too many bones to stage): bone ids are global, matrices come from L2.
// slots skinned per skin_vertex_n call: 2 measured best (1: -3.4 %, 4: spills, -11 %; profiles/r01_experiments.md)
constexpr int kSkinCallSlots = 2;
static_assert(kSlotGroup % kSkinCallSlots == 0, "a slot group is a whole number of skin calls");
template <int LAYOUT, bool EXT, bool PALG>
__global__ void __launch_bounds__(kSkinThreads, EXT ? 2 : 3) skin_kernel(DevModel M, DevFrames F, uint32_t chunk, uint32_t n_chunks) {
    constexpr uint32_t PS = EXT ? 5u : 3u;  // float4 per staged bone
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_bytes = skin_stage_bytes(LAYOUT, EXT), pal_bytes = skin_pal_bytes(M.max_tile_bones, EXT);
    const uint32_t pal4 = pal_bytes >> 4;  // float4 per staged slot palette
    unsigned char* stage_base = smem_raw;
    float4* pal_base = reinterpret_cast<float4*>(smem_raw + G * stage_bytes);
    float4* rate_base = pal_base + 2u * G * pal4;

    const uint32_t tile = blockIdx.x / n_chunks, ck = blockIdx.x - tile * n_chunks;
    const uint32_t s0 = ck * chunk;                      // chunk is a multiple of G: groups never straddle work items
    const uint32_t s1 = min(F.n_slots, s0 + chunk);
    if (s0 >= s1) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // vertices of this tile that exist (the last tile is padded): only those are stored, so a caller-owned output buffer
    // (mmdgpu_frames_bind_output) needs nv records per slot, not nv_pad
    const uint32_t tile_nv = min(kTileVerts, M.nv - tile * kTileVerts);

    // ---- the tile's static streams: read once, kept in registers for every slot of this work item
    const uint32_t v0 = tile * kTileVerts + tid * kVertsPerThread;
    float px[V], py[V], pz[V], nx[V], ny[V], nz[V], uu[V], vv[V];
    uint32_t ilo[V], ihi[V], orig[V];
    float4 wv[V];
    if (V == 4) {
        const float4 PX = __ldg(reinterpret_cast<const float4*>(M.px + v0)), PY = __ldg(reinterpret_cast<const float4*>(M.py + v0)),
                     PZ = __ldg(reinterpret_cast<const float4*>(M.pz + v0)), NX = __ldg(reinterpret_cast<const float4*>(M.nx + v0)),
                     NY = __ldg(reinterpret_cast<const float4*>(M.ny + v0)), NZ = __ldg(reinterpret_cast<const float4*>(M.nz + v0));
        const float a[6][4] = {{PX.x, PX.y, PX.z, PX.w}, {PY.x, PY.y, PY.z, PY.w}, {PZ.x, PZ.y, PZ.z, PZ.w},
                               {NX.x, NX.y, NX.z, NX.w}, {NY.x, NY.y, NY.z, NY.w}, {NZ.x, NZ.y, NZ.z, NZ.w}};
#pragma unroll
        for (int j = 0; j < V; ++j) { px[j] = a[0][j]; py[j] = a[1][j]; pz[j] = a[2][j]; nx[j] = a[3][j]; ny[j] = a[4][j]; nz[j] = a[5][j]; }
        const uint2 OR = __ldg(reinterpret_cast<const uint2*>(M.orig + v0));
        const uint32_t o[4] = {OR.x & 0xFFFFu, OR.x >> 16, OR.y & 0xFFFFu, OR.y >> 16};
#pragma unroll
        for (int j = 0; j < V; ++j) orig[j] = o[j];
    } else {
        const float2 PX = __ldg(reinterpret_cast<const float2*>(M.px + v0)), PY = __ldg(reinterpret_cast<const float2*>(M.py + v0)),
                     PZ = __ldg(reinterpret_cast<const float2*>(M.pz + v0)), NX = __ldg(reinterpret_cast<const float2*>(M.nx + v0)),
                     NY = __ldg(reinterpret_cast<const float2*>(M.ny + v0)), NZ = __ldg(reinterpret_cast<const float2*>(M.nz + v0));
        px[0] = PX.x; px[1] = PX.y; py[0] = PY.x; py[1] = PY.y; pz[0] = PZ.x; pz[1] = PZ.y;
        nx[0] = NX.x; nx[1] = NX.y; ny[0] = NY.x; ny[1] = NY.y; nz[0] = NZ.x; nz[1] = NZ.y;
        const uint32_t OR = __ldg(reinterpret_cast<const uint32_t*>(M.orig + v0));
        orig[0] = OR & 0xFFFFu; orig[1] = OR >> 16;
    }
#pragma unroll
    for (int j = 0; j < V; j += 2) {
        const uint4 I = __ldg(reinterpret_cast<const uint4*>(M.ids + v0 + j));
        ilo[j] = I.x; ihi[j] = I.y; ilo[j + 1] = I.z; ihi[j + 1] = I.w;
        uu[j] = vv[j] = uu[j + 1] = vv[j + 1] = 0.f;
        if (EXT) {  // extensions morph the UV every slot; the compat sokol32 path re-reads the static UV at staging time
            const float4 UV = __ldg(reinterpret_cast<const float4*>(M.uv + v0 + j));
            uu[j] = UV.x; vv[j] = UV.y; uu[j + 1] = UV.z; vv[j + 1] = UV.w;
        }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) wv[j] = __ldg(M.weights + v0 + j);
    // sliced-ELL group headers of this warp's steps (warp-uniform addresses: one broadcast load each)
    uint32_t ebase[V], erounds[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const uint2 h = __ldg(M.ell_hdr + tile * kTileGroups + j * kSkinWarps + warp);
        ebase[j] = h.x + lane;
        erounds[j] = h.y;
    }
    uint32_t uvbase[V], uvrounds[V];
    if (EXT) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const uint2 h = __ldg(M.uv_ell_hdr + tile * kTileGroups + j * kSkinWarps + warp);
            uvbase[j] = h.x + lane;
            uvrounds[j] = h.y;
        }
    }
    // tile-local palette of one slot group: item i = (slot f = i / npal4, float4 q = i % npal4); float4 q is component
    // q % PS of tile bone q / PS: a matrix column (0..2) or, with extensions, the rotation quaternion / dual part
    const uint32_t tb0 = __ldg(M.tile_bone_begin + tile);
    // global-palette models stage nothing: their bone ids index the slot's palette in global memory directly
    const uint32_t npal4 = PALG ? 0u : (__ldg(M.tile_bone_begin + tile + 1) - tb0) * PS;
    const uint32_t n_items = npal4 * G;
    const uint32_t npad = M.n_nodes_pad;               // float4 per rate block (one float4 = the G slots of a node)
    auto item_source = [&](uint32_t i) -> uint32_t {   // bits 31:30 = slot within the group, 29 = extension array
        const uint32_t f = i / npal4, q = i - f * npal4;
        const uint32_t bone = (uint32_t)__ldg(M.tile_bones + tb0 + q / PS), comp = q % PS;
        return (f << 30) | (comp < 3u ? bone * 3u + comp : (0x20000000u | (bone * 2u + comp - 3u)));
    };
    auto item_dest = [&](uint32_t i) -> uint32_t { const uint32_t f = i / npal4; return f * pal4 + (i - f * npal4); };
    auto item_fetch = [&](uint32_t group_slot0, uint32_t src) -> float4 {
        const uint32_t slot = min(group_slot0 + (src >> 30), F.n_slots - 1u);  // a partial last group re-reads its last slot
        const uint32_t idx = src & 0x1FFFFFFFu;
        // (plain loads, not the read-only path: under programmatic dependent launch the hierarchy may still have been
        // writing these while this kernel was already resident)
        if (EXT && (src & 0x20000000u)) return F.pal_ext[(size_t)slot * M.nb * 2 + idx];
        return F.palette[(size_t)slot * M.nb * 3 + idx];
    };
    uint32_t psrc[kPalPrefetch], pdst[kPalPrefetch];
#pragma unroll
    for (uint32_t q = 0; q < kPalPrefetch; ++q) {
        const uint32_t i = tid + q * kSkinThreads;
        psrc[q] = (i < n_items) ? item_source(i) : 0xFFFFFFFFu;
        pdst[q] = (i < n_items) ? item_dest(i) : 0u;
    }

    // ---- prologue: first group straight into buffer 0 (slot state: only now does a dependent launch wait for the hierarchy)
    pdl_wait();
    {
        for (uint32_t i = tid; i < n_items; i += kSkinThreads) pal_base[item_dest(i)] = item_fetch(s0, item_source(i));
        const float4* gr = reinterpret_cast<const float4*>(F.node_rate) + (size_t)(s0 / G) * npad;
        for (uint32_t i = tid; i < npad; i += kSkinThreads) rate_base[i] = gr[i];
    }
    __syncthreads();

    uint32_t b = 0;
    for (uint32_t g0 = s0; g0 < s1; g0 += G, b ^= 1u) {
        const float4* __restrict__ pal = pal_base + (size_t)b * G * pal4;
        const char* __restrict__ nrate = reinterpret_cast<const char*>(rate_base + (size_t)b * npad);
        const uint32_t n_live = min((uint32_t)G, s1 - g0);   // slots of this group that exist (CTA-uniform)
        const bool has_next = g0 + G < s1;
        // ---- next group's palettes and rates: loads issued now, consumed after the compute phase
        float4 pf[kPalPrefetch], rf = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* gr = reinterpret_cast<const float4*>(F.node_rate) + (size_t)(g0 / G + 1) * npad;
        if (has_next) {
#pragma unroll
            for (uint32_t q = 0; q < kPalPrefetch; ++q)
                if (psrc[q] != 0xFFFFFFFFu) pf[q] = item_fetch(g0 + G, psrc[q]);
            if (tid < npad) rf = gr[tid];
        }
        // ---- per storage position: morph gather for the G slots at once, then G skinnings.  vertex_images_[i]
        //      accumulates in application order: img = img + off*rate (poser_impl.inl:340-346).  ent.w = byte
        //      offset of the entry's float4 of rates (one per slot of the group).  A skipped slot has rate +0 and
        //      is applied unconditionally: the accumulator starts at +0 and can never become -0, so adding
        //      (finite offset) * 0 = +-0 leaves it bit-identical to libmmd's skip (the host rejects non-finite offsets).
        //      Step j is (nearly always) one skinning type across the warp.
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float ix[G], iy[G], iz[G];
#pragma unroll
            for (int f = 0; f < G; ++f) ix[f] = iy[f] = iz[f] = 0.f;
            const float4* __restrict__ e = M.ell_ent + ebase[j];
            auto accumulate = [&](const float4& ent) {
                const float4 r4 = *reinterpret_cast<const float4*>(nrate + __float_as_uint(ent.w));
                const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int f = 0; f < G; ++f) {
                    ix[f] = ix[f] + ent.x * r[f];
                    iy[f] = iy[f] + ent.y * r[f];
                    iz[f] = iz[f] + ent.z * r[f];
                }
            };
            for (uint32_t k = 0; k < erounds[j]; ++k) accumulate(__ldg(e + (size_t)k * 32));
            if (j == 0 && LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) {
                // the staging tiles are single-buffered: the previous group's bulk copies must have read them
                if (tid == 0) bulk_wait_read_all();
                __syncthreads();
            }
            const uint32_t type = (ilo[j] >> 13) & 7u;
            float su = uu[j], sv_ = vv[j];
            if (LAYOUT == MMDGPU_LAYOUT_INTERLEAVED_SOKOL32 && !EXT) {
                // static UV passthrough (main.cpp:840,855-856): an L1-resident 8-byte load, no registers held across slots
                const float2 t = __ldg(M.uv + v0 + j);
                su = t.x; sv_ = t.y;
            }
            if (!EXT && !PALG) {
                constexpr int NS = kSkinCallSlots;
#pragma unroll
                for (int f = 0; f < G; f += NS) {
                    MorphedN<NS> q;
#pragma unroll
                    for (int h = 0; h < NS; ++h) { q.x[h] = px[j] + ix[f + h]; q.y[h] = py[j] + iy[f + h]; q.z[h] = pz[j] + iz[f + h]; }
                    const SkinnedN<NS> rn = skin_vertex_n<(int)PS, true, NS>(pal + (size_t)f * pal4, pal4, ilo[j], ihi[j], wv[j], q,
                                                                             nx[j], ny[j], nz[j]);
#pragma unroll
                    for (int h = 0; h < NS; ++h) {
                        if ((uint32_t)(f + h) >= n_live) continue;
                        const Skinned& r = rn.s[h];
                        if (LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) {
                            unsigned char* stage = stage_base + (size_t)(f + h) * stage_bytes;
                            float* sp = reinterpret_cast<float*>(stage) + orig[j] * 3u;
                            float* sn = reinterpret_cast<float*>(stage + kTileVerts * 12u) + orig[j] * 3u;
                            sp[0] = r.px; sp[1] = r.py; sp[2] = r.pz;
                            sn[0] = r.nx; sn[1] = r.ny; sn[2] = r.nz;
                        } else if (orig[j] < tile_nv) {
                            const float mmd_to_meter = 0.1f;
                            float4* sv = F.out_inter + (size_t)(g0 + f + h) * F.inter_stride + ((size_t)tile * kTileVerts + orig[j]) * 2u;
                            store_record32(sv, r.px * mmd_to_meter, r.py * mmd_to_meter, r.pz * mmd_to_meter, r.nx, r.ny, r.nz, su, sv_);
                        }
                    }
                }
            } else {
#pragma unroll
            for (int f = 0; f < G; ++f) {
                // slots past the end of a partial last group are computed on the clamped palette and not stored
                const bool live = (uint32_t)f < n_live;
                const float4* __restrict__ palf =
                    PALG ? F.palette + (size_t)min(g0 + (uint32_t)f, F.n_slots - 1u) * M.nb * 3 : pal + (size_t)f * pal4;
                unsigned char* stage = stage_base + (size_t)f * stage_bytes;
                float op[3], on[3];
                // coordinate + vertex_image (poser_impl.inl:407)
                const float qx = px[j] + ix[f], qy = py[j] + iy[f], qz = pz[j] + iz[f];
                if (EXT && type == kDevSdef)
                    skin_sdef(palf, ilo[j] & 0x1FFFu, ilo[j] >> 16, wv[j].x, M.sdef + (size_t)(v0 + j) * 3, qx, qy, qz, nx[j], ny[j],
                              nz[j], op, on);
                else if (EXT && type == kDevQdef)
                    skin_qdef(palf, ilo[j], ihi[j], wv[j], qx, qy, qz, nx[j], ny[j], nz[j], op, on);
                else {
                    const Skinned r = skin_vertex<(int)PS, !PALG>(palf, ilo[j], ihi[j], wv[j], qx, qy, qz, nx[j], ny[j], nz[j]);
                    op[0] = r.px; op[1] = r.py; op[2] = r.pz;
                    on[0] = r.nx; on[1] = r.ny; on[2] = r.nz;
                }
                float mu = su, mv = sv_;
                if (EXT) {
                    // applied UV morphs: uv = uv + offset.xy * rate, application order
                    const float4* __restrict__ ue = M.uv_ell_ent + uvbase[j];
                    for (uint32_t k = 0; k < uvrounds[j]; ++k) {
                        const float4 ent = __ldg(ue + (size_t)k * 32);
                        const float r = *reinterpret_cast<const float*>(nrate + __float_as_uint(ent.z) + 4u * f);
                        mu = mu + ent.x * r;
                        mv = mv + ent.y * r;
                    }
                }
                if (!live) continue;
                if (LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) {
                    float* sp = reinterpret_cast<float*>(stage) + orig[j] * 3u;
                    float* sn = reinterpret_cast<float*>(stage + kTileVerts * 12u) + orig[j] * 3u;
                    sp[0] = op[0]; sp[1] = op[1]; sp[2] = op[2];
                    sn[0] = on[0]; sn[1] = on[1]; sn[2] = on[2];
                    if (EXT) reinterpret_cast<float2*>(stage + kTileVerts * 24u)[orig[j]] = make_float2(mu, mv);
                } else if (orig[j] < tile_nv) {
                    // main.cpp:838-859: Vertex{pos*0.1f, normal, uv}
                    const float mmd_to_meter = 0.1f;
                    float4* sv = F.out_inter + (size_t)(g0 + f) * F.inter_stride + ((size_t)tile * kTileVerts + orig[j]) * 2u;
                    store_record32(sv, op[0] * mmd_to_meter, op[1] * mmd_to_meter, op[2] * mmd_to_meter, on[0], on[1], on[2], mu, mv);
                }
            }
            }
        }
        // ---- publish the next group's palettes / rates into the other buffer
        if (has_next) {
            float4* npal = pal_base + (size_t)(b ^ 1u) * G * pal4;
            float4* nrt = rate_base + (size_t)(b ^ 1u) * npad;
#pragma unroll
            for (uint32_t q = 0; q < kPalPrefetch; ++q)
                if (psrc[q] != 0xFFFFFFFFu) npal[pdst[q]] = pf[q];
            for (uint32_t i = tid + kPalPrefetch * kSkinThreads; i < n_items; i += kSkinThreads)
                npal[item_dest(i)] = item_fetch(g0 + G, item_source(i));
            if (tid < npad) nrt[tid] = rf;
            for (uint32_t i = tid + kSkinThreads; i < npad; i += kSkinThreads) nrt[i] = gr[i];
        }
        // ---- hand the staged tiles to the bulk-copy engine (the barrier also orders the palette double buffer)
        if (LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) fence_proxy_async_smem();  // staging writes -> visible to the async proxy
        __syncthreads();
        if (tid == 0 && LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) {
            // bulk copies move whole 16-byte units; a last tile whose vertex count is not a multiple of 4 leaves up to
            // three floats per plane, stored by this thread
            const uint32_t b3 = (tile_nv * 12u) & ~15u, b2 = (tile_nv * 8u) & ~15u;
            for (uint32_t f = 0; f < n_live; ++f) {
                unsigned char* stage = stage_base + (size_t)f * stage_bytes;
                const size_t vt = (size_t)tile * kTileVerts;
                float* dp = F.out_pos + (size_t)(g0 + f) * F.pos_stride + vt * 3;
                float* dn = F.out_nrm + (size_t)(g0 + f) * F.nrm_stride + vt * 3;
                if (b3) {
                    bulk_s2g(dp, stage, b3);
                    bulk_s2g(dn, stage + kTileVerts * 12u, b3);
                }
                for (uint32_t w = b3 / 4u; w < tile_nv * 3u; ++w) {
                    dp[w] = reinterpret_cast<const float*>(stage)[w];
                    dn[w] = reinterpret_cast<const float*>(stage + kTileVerts * 12u)[w];
                }
                if (EXT) {
                    float2* du = F.out_uv + (size_t)(g0 + f) * F.uv_stride + vt;
                    if (b2) bulk_s2g(du, stage + kTileVerts * 24u, b2);
                    for (uint32_t w = b2 / 8u; w < tile_nv; ++w) du[w] = reinterpret_cast<const float2*>(stage + kTileVerts * 24u)[w];
                }
            }
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all();  // shared memory must outlive the copies that read it
}

// =================================================================================================
// K3, packed form (the kernel every libmmd-exact model with staged tile palettes runs).
//
// Same work decomposition as skin_kernel above, but all per-slot arithmetic is done for TWO slots at once in packed
// fp32x2 registers (Blackwell FFMA2): the blend and the transform of slots (2p, 2p+1) of one vertex are the same
// operations on different data, so one issue slot carries both.  Bit parity with libmmd needs un-fused multiplies and
// adds; sm_100 has no packed FMUL / FADD, and ptxas contracts `mul.f32x2` + `add.f32x2` into one FFMA2 even when both
// carry .rn, so a product is written  fma(a, b, -0)  and a sum  fma(a, 1, b)  with -0 and 1 coming from kernel
// ARGUMENTS (opaque to the compiler).  Both are exact restatements: a*b + (-0) rounds the exact product once and keeps
// its zero sign, a*1 + b is a + b (tools/micro/ffma2.cu checks 4 M operand triples incl. zeros, denormals, infinities).
// For this the staged palettes are stored pair-interleaved: per slot pair two planes of 16-byte cells,
// lo[bone][col] = (x0, x1, y0, y1) and hi[bone][col] = (z0, z1, w0, w1), so that one LDS.128 yields two ready pairs.
// The four slots' morph rates of a node are already one float4 = two pairs.
// =================================================================================================
typedef unsigned long long f2;   // two floats in one 64-bit register: (lo, hi) = (even slot, odd slot)
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
struct PairK { f2 nz, one; };    // (-0, -0) and (1, 1), from kernel arguments
__device__ __forceinline__ f2 mul2(f2 a, f2 b, const PairK& K) { return fma2(a, b, K.nz); }
__device__ __forceinline__ f2 add2(f2 a, f2 b, const PairK& K) { return fma2(a, K.one, b); }

struct SkinnedP { f2 px, py, pz, nx, ny, nz; };
struct ColP { f2 x, y, z, w; };  // one matrix column (M0c, M1c, M2c, M3c) of a slot pair
__device__ __forceinline__ ColP colp_load(const ulonglong2* __restrict__ pal, uint32_t hi_off, uint32_t cell) {
    const ulonglong2 lo = pal[cell], hi = pal[hi_off + cell];
    return ColP{lo.x, lo.y, hi.x, hi.y};
}
// One vertex, two slots.  pal: the slot pair's lo plane; hi_off: distance to its hi plane (16-byte cells).
// Poser::Deform's blend (poser_impl.inl:404-436) and transform / rotate (math_impl.inl:1032-1045), association order kept.
__device__ __noinline__ SkinnedP skin_vertex_pair(const ulonglong2* __restrict__ pal, uint32_t hi_off, uint32_t ids_lo, uint32_t ids_hi,
                                                  float4 w, f2 qx, f2 qy, f2 qz, float nx, float ny, float nz, f2 k_nz, f2 k_one) {
    __builtin_assume(__isShared(pal));
    const PairK K{k_nz, k_one};
    const uint32_t type = (ids_lo >> 13) & 7u;
    const uint32_t id0 = ids_lo & 0x1FFFu, id1 = ids_lo >> 16;
    ColP m[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) m[c] = colp_load(pal, hi_off, id0 * 3u + c);
    if (type == kDevBdef2) {
        // Lerp(mat_1, mat_0)[w] = (1-w)*mat_1 + w*mat_0
        const float om1 = 1.0f - w.x;
        const f2 L = pk2(w.x, w.x), OM = pk2(om1, om1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const ColP b = colp_load(pal, hi_off, id1 * 3u + c);
            m[c].x = add2(mul2(b.x, OM, K), mul2(m[c].x, L, K), K);
            m[c].y = add2(mul2(b.y, OM, K), mul2(m[c].y, L, K), K);
            m[c].z = add2(mul2(b.z, OM, K), mul2(m[c].z, L, K), K);
            m[c].w = add2(mul2(b.w, OM, K), mul2(m[c].w, L, K), K);
        }
    } else if (type == kDevBdef4) {
        // mat_0*w0 + mat_1*w1 + mat_2*w2 + mat_3*w3, left to right
        const uint32_t id2 = ids_hi & 0xFFFFu, id3 = ids_hi >> 16;
        const f2 W0 = pk2(w.x, w.x), W1 = pk2(w.y, w.y), W2 = pk2(w.z, w.z), W3 = pk2(w.w, w.w);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const ColP b = colp_load(pal, hi_off, id1 * 3u + c), cc = colp_load(pal, hi_off, id2 * 3u + c), d = colp_load(pal, hi_off, id3 * 3u + c);
            m[c].x = add2(add2(add2(mul2(m[c].x, W0, K), mul2(b.x, W1, K), K), mul2(cc.x, W2, K), K), mul2(d.x, W3, K), K);
            m[c].y = add2(add2(add2(mul2(m[c].y, W0, K), mul2(b.y, W1, K), K), mul2(cc.y, W2, K), K), mul2(d.y, W3, K), K);
            m[c].z = add2(add2(add2(mul2(m[c].z, W0, K), mul2(b.z, W1, K), K), mul2(cc.z, W2, K), K), mul2(d.z, W3, K), K);
            m[c].w = add2(add2(add2(mul2(m[c].w, W0, K), mul2(b.w, W1, K), K), mul2(cc.w, W2, K), K), mul2(d.w, W3, K), K);
        }
    }
    const f2 NX = pk2(nx, nx), NY = pk2(ny, ny), NZ = pk2(nz, nz);
    SkinnedP r;
    r.px = add2(add2(add2(mul2(qx, m[0].x, K), mul2(qy, m[0].y, K), K), mul2(qz, m[0].z, K), K), m[0].w, K);
    r.py = add2(add2(add2(mul2(qx, m[1].x, K), mul2(qy, m[1].y, K), K), mul2(qz, m[1].z, K), K), m[1].w, K);
    r.pz = add2(add2(add2(mul2(qx, m[2].x, K), mul2(qy, m[2].y, K), K), mul2(qz, m[2].z, K), K), m[2].w, K);
    r.nx = add2(add2(mul2(NX, m[0].x, K), mul2(NY, m[0].y, K), K), mul2(NZ, m[0].z, K), K);
    r.ny = add2(add2(mul2(NX, m[1].x, K), mul2(NY, m[1].y, K), K), mul2(NZ, m[1].z, K), K);
    r.nz = add2(add2(mul2(NX, m[2].x, K), mul2(NY, m[2].y, K), K), mul2(NZ, m[2].z, K), K);
    return r;
}

template <int LAYOUT, int MIN_CTAS, bool STAGED>
__global__ void __launch_bounds__(kSkinThreads, MIN_CTAS) skin_pair_kernel(DevModel M, DevFrames F, uint32_t chunk, uint32_t n_chunks,
                                                                    float arg_neg_zero, float arg_one) {
    static_assert(STAGED || LAYOUT != MMDGPU_LAYOUT_SOA_POS_NRM, "the planar layout always leaves through staging tiles");
    static_assert(G == 4, "two slot pairs per group");
    constexpr int NP = G / 2;                           // slot pairs per group
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool staged = STAGED;
    constexpr uint32_t stage_bytes = skin_pair_stage_bytes(LAYOUT, staged);
    const uint32_t pal_bytes = skin_pal_bytes(M.max_tile_bones, false);
    const uint32_t pal4 = pal_bytes >> 4;               // 16-byte cells per slot; a pair owns 2 * pal4: lo plane, hi plane
    unsigned char* stage_base = smem_raw;
    ulonglong2* pal_base = reinterpret_cast<ulonglong2*>(smem_raw + G * stage_bytes);
    float4* rate_base = reinterpret_cast<float4*>(pal_base + 2u * G * pal4);
    const PairK K{pk2(arg_neg_zero, arg_neg_zero), pk2(arg_one, arg_one)};

    const uint32_t tile = blockIdx.x / n_chunks, ck = blockIdx.x - tile * n_chunks;
    const uint32_t s0 = ck * chunk;
    const uint32_t s1 = min(F.n_slots, s0 + chunk);
    if (s0 >= s1) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile_nv = min(kTileVerts, M.nv - tile * kTileVerts);

    // ---- the tile's static streams: read once, kept in registers for every slot of this work item
    const uint32_t v0 = tile * kTileVerts + tid * kVertsPerThread;
    float px[V], py[V], pz[V], nx[V], ny[V], nz[V];
    uint32_t ilo[V], ihi[V], orig[V];
    float4 wv[V];
    if (V == 4) {
        const float4 PX = __ldg(reinterpret_cast<const float4*>(M.px + v0)), PY = __ldg(reinterpret_cast<const float4*>(M.py + v0)),
                     PZ = __ldg(reinterpret_cast<const float4*>(M.pz + v0)), NX = __ldg(reinterpret_cast<const float4*>(M.nx + v0)),
                     NY = __ldg(reinterpret_cast<const float4*>(M.ny + v0)), NZ = __ldg(reinterpret_cast<const float4*>(M.nz + v0));
        const float a[6][4] = {{PX.x, PX.y, PX.z, PX.w}, {PY.x, PY.y, PY.z, PY.w}, {PZ.x, PZ.y, PZ.z, PZ.w},
                               {NX.x, NX.y, NX.z, NX.w}, {NY.x, NY.y, NY.z, NY.w}, {NZ.x, NZ.y, NZ.z, NZ.w}};
#pragma unroll
        for (int j = 0; j < V; ++j) { px[j] = a[0][j]; py[j] = a[1][j]; pz[j] = a[2][j]; nx[j] = a[3][j]; ny[j] = a[4][j]; nz[j] = a[5][j]; }
        const uint2 OR = __ldg(reinterpret_cast<const uint2*>(M.orig + v0));
        const uint32_t o[4] = {OR.x & 0xFFFFu, OR.x >> 16, OR.y & 0xFFFFu, OR.y >> 16};
#pragma unroll
        for (int j = 0; j < V; ++j) orig[j] = o[j];
    } else {
        const float2 PX = __ldg(reinterpret_cast<const float2*>(M.px + v0)), PY = __ldg(reinterpret_cast<const float2*>(M.py + v0)),
                     PZ = __ldg(reinterpret_cast<const float2*>(M.pz + v0)), NX = __ldg(reinterpret_cast<const float2*>(M.nx + v0)),
                     NY = __ldg(reinterpret_cast<const float2*>(M.ny + v0)), NZ = __ldg(reinterpret_cast<const float2*>(M.nz + v0));
        px[0] = PX.x; px[1] = PX.y; py[0] = PY.x; py[1] = PY.y; pz[0] = PZ.x; pz[1] = PZ.y;
        nx[0] = NX.x; nx[1] = NX.y; ny[0] = NY.x; ny[1] = NY.y; nz[0] = NZ.x; nz[1] = NZ.y;
        const uint32_t OR = __ldg(reinterpret_cast<const uint32_t*>(M.orig + v0));
        orig[0] = OR & 0xFFFFu; orig[1] = OR >> 16;
    }
#pragma unroll
    for (int j = 0; j < V; j += 2) {
        const uint4 I = __ldg(reinterpret_cast<const uint4*>(M.ids + v0 + j));
        ilo[j] = I.x; ihi[j] = I.y; ilo[j + 1] = I.z; ihi[j + 1] = I.w;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) wv[j] = __ldg(M.weights + v0 + j);
    uint32_t ebase[V], erounds[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const uint2 h = __ldg(M.ell_hdr + tile * kTileGroups + j * kSkinWarps + warp);
        ebase[j] = h.x + lane;
        erounds[j] = h.y;
    }
    // ---- tile-local palettes of one slot group, pair-interleaved.  Item i = (pair p = i / npal4, cell q = i % npal4); cell q
    //      is column q % 3 of tile bone q / 3.  Its two source float4 (slots 2p and 2p+1 of the group) become the cells
    //      lo = (x0, x1, y0, y1) and hi = (z0, z1, w0, w1).
    const uint32_t tb0 = __ldg(M.tile_bone_begin + tile);
    const uint32_t npal4 = (__ldg(M.tile_bone_begin + tile + 1) - tb0) * 3u;
    const uint32_t n_items = npal4 * NP;
    const uint32_t npad = M.n_nodes_pad;
    auto item_cell = [&](uint32_t i, uint32_t& pair, uint32_t& q) -> uint32_t {   // returns the source float4 index inside a slot's palette
        pair = i / npal4; q = i - pair * npal4;
        return (uint32_t)__ldg(M.tile_bones + tb0 + q / 3u) * 3u + q % 3u;
    };
    auto fetch = [&](uint32_t group_slot0, uint32_t pair, uint32_t src, float4& A, float4& B) {
        const uint32_t sa = min(group_slot0 + 2u * pair, F.n_slots - 1u), sb = min(group_slot0 + 2u * pair + 1u, F.n_slots - 1u);
        // (plain loads, not the read-only path: under programmatic dependent launch the hierarchy may still have been
        // writing these while this kernel was already resident)
        A = F.palette[(size_t)sa * M.nb * 3 + src];
        B = F.palette[(size_t)sb * M.nb * 3 + src];
    };
    auto publish = [&](ulonglong2* buf, uint32_t pair, uint32_t q, const float4& A, const float4& B) {
        ulonglong2* cell = buf + (size_t)pair * 2u * pal4 + q;
        cell[0] = make_ulonglong2(pk2(A.x, B.x), pk2(A.y, B.y));
        cell[pal4] = make_ulonglong2(pk2(A.z, B.z), pk2(A.w, B.w));
    };
    // one item per thread is prefetched in registers across the compute phase (C3: 2 pairs x ~8 bones x 3 = 48 items)
    uint32_t my_pair = 0, my_q = 0, my_src = 0xFFFFFFFFu;
    if (tid < n_items) my_src = item_cell(tid, my_pair, my_q);

    // ---- prologue: first group straight into buffer 0 (slot state: only now does a dependent launch wait for the hierarchy)
    pdl_wait();
    {
        for (uint32_t i = tid; i < n_items; i += kSkinThreads) {
            uint32_t p, q; float4 A, B;
            const uint32_t src = item_cell(i, p, q);
            fetch(s0, p, src, A, B);
            publish(pal_base, p, q, A, B);
        }
        const float4* gr = reinterpret_cast<const float4*>(F.node_rate) + (size_t)(s0 / G) * npad;
        for (uint32_t i = tid; i < npad; i += kSkinThreads) rate_base[i] = gr[i];
    }
    __syncthreads();

    uint32_t b = 0;
    for (uint32_t g0 = s0; g0 < s1; g0 += G, b ^= 1u) {
        const ulonglong2* __restrict__ pal = pal_base + (size_t)b * G * pal4;
        const char* __restrict__ nrate = reinterpret_cast<const char*>(rate_base + (size_t)b * npad);
        const uint32_t n_live = min((uint32_t)G, s1 - g0);
        const bool has_next = g0 + G < s1;
        // ---- next group's palettes and rates: loads issued now, consumed after the compute phase
        float4 pfA = make_float4(0.f, 0.f, 0.f, 0.f), pfB = pfA, rf = pfA;
        const float4* gr = reinterpret_cast<const float4*>(F.node_rate) + (size_t)(g0 / G + 1) * npad;
        if (has_next) {
            if (my_src != 0xFFFFFFFFu) fetch(g0 + G, my_pair, my_src, pfA, pfB);
            if (tid < npad) rf = gr[tid];
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            // vertex_images_[i] of the four slots as two pairs; img = img + off * rate in application order (poser_impl.inl:340-346).
            // A skipped slot has rate +0 and is applied unconditionally (the accumulator starts at +0 and can never become -0).
            f2 IX[NP], IY[NP], IZ[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) IX[p] = IY[p] = IZ[p] = 0ull;
            const float4* __restrict__ e = M.ell_ent + ebase[j];
            for (uint32_t k = 0; k < erounds[j]; ++k) {
                const float4 ent = __ldg(e + (size_t)k * 32);
                const ulonglong2 r = *reinterpret_cast<const ulonglong2*>(nrate + __float_as_uint(ent.w));
                const f2 EX = pk2(ent.x, ent.x), EY = pk2(ent.y, ent.y), EZ = pk2(ent.z, ent.z);
                IX[0] = add2(IX[0], mul2(EX, r.x, K), K); IX[1] = add2(IX[1], mul2(EX, r.y, K), K);
                IY[0] = add2(IY[0], mul2(EY, r.x, K), K); IY[1] = add2(IY[1], mul2(EY, r.y, K), K);
                IZ[0] = add2(IZ[0], mul2(EZ, r.x, K), K); IZ[1] = add2(IZ[1], mul2(EZ, r.y, K), K);
            }
            if (j == 0 && staged) {
                // the staging tiles are single-buffered: the previous group's bulk copies must have read them
                if (tid == 0) bulk_wait_read_all();
                __syncthreads();
            }
            float su = 0.f, sv_ = 0.f;
            if (LAYOUT == MMDGPU_LAYOUT_INTERLEAVED_SOKOL32) {
                const float2 t = __ldg(M.uv + v0 + j);   // static UV passthrough (main.cpp:840,855-856)
                su = t.x; sv_ = t.y;
            }
            const f2 PX2 = pk2(px[j], px[j]), PY2 = pk2(py[j], py[j]), PZ2 = pk2(pz[j], pz[j]);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if ((uint32_t)(2 * p) >= n_live) continue;   // a partial last group (a one-slot Poser): nothing to store for this pair
                // coordinate + vertex_image (poser_impl.inl:407)
                const SkinnedP r = skin_vertex_pair(pal + (size_t)p * 2u * pal4, pal4, ilo[j], ihi[j], wv[j], add2(PX2, IX[p], K),
                                                    add2(PY2, IY[p], K), add2(PZ2, IZ[p], K), nx[j], ny[j], nz[j], K.nz, K.one);
                if (LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if ((uint32_t)(2 * p + h) >= n_live) continue;
                        unsigned char* stage = stage_base + (size_t)(2 * p + h) * stage_bytes;
                        float* sp = reinterpret_cast<float*>(stage) + orig[j] * 3u;
                        float* sn = reinterpret_cast<float*>(stage + kTileVerts * 12u) + orig[j] * 3u;
                        sp[0] = h ? hi2(r.px) : lo2(r.px); sp[1] = h ? hi2(r.py) : lo2(r.py); sp[2] = h ? hi2(r.pz) : lo2(r.pz);
                        sn[0] = h ? hi2(r.nx) : lo2(r.nx); sn[1] = h ? hi2(r.ny) : lo2(r.ny); sn[2] = h ? hi2(r.nz) : lo2(r.nz);
                    }
                } else if (staged) {
                    // main.cpp:838-859: Vertex{pos * 0.1f, normal, uv}, staged at 32 bytes x PMX index as two 16-byte halves.
                    // A lane whose index has bit 2 set stores its halves in the opposite order: the 16-byte bank group of a
                    // store is (2 * index + half) mod 8, so 32 lanes with distinct indices mod 32 cover all eight groups
                    // evenly in each of the two STS.128 (4 wavefronts each = the floor for 512 bytes).
                    const f2 T = pk2(0.1f, 0.1f);
                    const f2 sx = mul2(r.px, T, K), sy = mul2(r.py, T, K), sz = mul2(r.pz, T, K);
                    const bool flip = (orig[j] & 4u) != 0u;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if ((uint32_t)(2 * p + h) >= n_live) continue;
                        float4* rec = reinterpret_cast<float4*>(stage_base + (size_t)(2 * p + h) * stage_bytes) + orig[j] * 2u;
                        const float4 lo = make_float4(h ? hi2(sx) : lo2(sx), h ? hi2(sy) : lo2(sy), h ? hi2(sz) : lo2(sz), h ? hi2(r.nx) : lo2(r.nx));
                        const float4 hi = make_float4(h ? hi2(r.ny) : lo2(r.ny), h ? hi2(r.nz) : lo2(r.nz), su, sv_);
                        rec[flip ? 1 : 0] = flip ? hi : lo;
                        rec[flip ? 0 : 1] = flip ? lo : hi;
                    }
                } else if (orig[j] < tile_nv) {
                    const f2 T = pk2(0.1f, 0.1f);
                    const f2 sx = mul2(r.px, T, K), sy = mul2(r.py, T, K), sz = mul2(r.pz, T, K);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if ((uint32_t)(2 * p + h) >= n_live) continue;
                        float4* sv = F.out_inter + (size_t)(g0 + 2 * p + h) * F.inter_stride + ((size_t)tile * kTileVerts + orig[j]) * 2u;
                        store_record32(sv, h ? hi2(sx) : lo2(sx), h ? hi2(sy) : lo2(sy), h ? hi2(sz) : lo2(sz), h ? hi2(r.nx) : lo2(r.nx),
                                       h ? hi2(r.ny) : lo2(r.ny), h ? hi2(r.nz) : lo2(r.nz), su, sv_);
                    }
                }
            }
        }
        // ---- publish the next group's palettes / rates into the other buffer
        if (has_next) {
            ulonglong2* npal = pal_base + (size_t)(b ^ 1u) * G * pal4;
            float4* nrt = rate_base + (size_t)(b ^ 1u) * npad;
            if (my_src != 0xFFFFFFFFu) publish(npal, my_pair, my_q, pfA, pfB);
            for (uint32_t i = tid + kSkinThreads; i < n_items; i += kSkinThreads) {
                uint32_t p, q; float4 A, B;
                const uint32_t src = item_cell(i, p, q);
                fetch(g0 + G, p, src, A, B);
                publish(npal, p, q, A, B);
            }
            if (tid < npad) nrt[tid] = rf;
            for (uint32_t i = tid + kSkinThreads; i < npad; i += kSkinThreads) nrt[i] = gr[i];
        }
        // ---- hand the staged tiles to the bulk-copy engine (the barrier also orders the palette double buffer)
        if (staged) fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && LAYOUT == MMDGPU_LAYOUT_SOA_POS_NRM) {
            const uint32_t b3 = (tile_nv * 12u) & ~15u;
            for (uint32_t f = 0; f < n_live; ++f) {
                unsigned char* stage = stage_base + (size_t)f * stage_bytes;
                const size_t vt = (size_t)tile * kTileVerts;
                float* dp = F.out_pos + (size_t)(g0 + f) * F.pos_stride + vt * 3;
                float* dn = F.out_nrm + (size_t)(g0 + f) * F.nrm_stride + vt * 3;
                if (b3) {
                    bulk_s2g(dp, stage, b3);
                    bulk_s2g(dn, stage + kTileVerts * 12u, b3);
                }
                for (uint32_t w = b3 / 4u; w < tile_nv * 3u; ++w) {
                    dp[w] = reinterpret_cast<const float*>(stage)[w];
                    dn[w] = reinterpret_cast<const float*>(stage + kTileVerts * 12u)[w];
                }
            }
            bulk_commit();
        } else if (tid == 0 && staged) {
            // one copy per slot: the tile's records are contiguous and whole multiples of 16 bytes
            for (uint32_t f = 0; f < n_live; ++f)
                bulk_s2g(F.out_inter + (size_t)(g0 + f) * F.inter_stride + (size_t)tile * kTileVerts * 2u,
                         stage_base + (size_t)f * stage_bytes, tile_nv * 32u);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all();  // shared memory must outlive the copies that read it
}

// =================================================================================================
// Function-level known-answer kernel (mmdgpu_test_math): one thread per case runs ONE device function of
// mmd_math.cuh on a row of inputs.  Row layouts: oracle/mmd_oracle.c, port_math_kat.
// =================================================================================================
__global__ void math_kat_kernel(int op, const float* __restrict__ in, uint32_t n, float* __restrict__ out,
                                const float* __restrict__ tables, const uint32_t* __restrict__ curve) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int kin[11] = {5, 9, 9, 5, 4, 4, 8, 4, 4, 32, 3}, kout[11] = {1, 4, 4, 3, 4, 4, 4, 9, 4, 16, 3};
    const float* a = in + (size_t)i * kin[op];
    float* o = out + (size_t)i * kout[op];
    auto putq = [&](const Quat& q) { o[0] = q.i; o[1] = q.j; o[2] = q.k; o[3] = q.e; };
    switch (op) {
    case 0: o[0] = bezier_at(tables, curve[i], a[4]); break;   // table built on the host by bezier_table(), looked up here
    case 1: {
        const float4 v = v4_nlerp(make_float4(a[0], a[1], a[2], a[3]), make_float4(a[4], a[5], a[6], a[7]), a[8]);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        break;
    }
    case 2: putq(q_slerp(Quat{a[0], a[1], a[2], a[3]}, Quat{a[4], a[5], a[6], a[7]}, a[8])); break;
    case 3: { const Vec3 e = quat_to_euler((int)a[4], Quat{a[0], a[1], a[2], a[3]}); o[0] = e.x; o[1] = e.y; o[2] = e.z; break; }
    case 4: putq(euler_to_quat((int)a[3], Vec3{a[0], a[1], a[2]})); break;
    case 5: putq(axis_to_quat(Vec3{a[0], a[1], a[2]}, a[3])); break;
    case 6: putq(q_mul(Quat{a[0], a[1], a[2], a[3]}, Quat{a[4], a[5], a[6], a[7]})); break;
    case 7: {
        Mat43 M;
        q_to_rows(Quat{a[0], a[1], a[2], a[3]}, M);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = M.m[r][c];
        break;
    }
    case 8: putq(q_inverse(Quat{a[0], a[1], a[2], a[3]})); break;
    case 9: {   // affine 4 x 4 operands only (fourth column 0,0,0,1): the device keeps 12 elements (mmd_math.cuh, m_mul)
        Mat43 A, B;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 3; ++c) { A.m[r][c] = a[4 * r + c]; B.m[r][c] = a[16 + 4 * r + c]; }
        const Mat43 R = m_mul(A, B);
        for (int r = 0; r < 4; ++r) { for (int c = 0; c < 3; ++c) o[4 * r + c] = R.m[r][c]; o[4 * r + 3] = r == 3 ? 1.0f : 0.0f; }
        break;
    }
    case 10: { const Vec3 v = v_normalize(Vec3{a[0], a[1], a[2]}); o[0] = v.x; o[1] = v.y; o[2] = v.z; break; }
    default: break;
    }
}

cudaError_t launch_math_kat(cudaStream_t st, int op, const float* in, uint32_t n, float* out, const float* tables, const uint32_t* curve) {
    if (n == 0) return cudaSuccess;
    math_kat_kernel<<<(n + 127) / 128, 128, 0, st>>>(op, in, n, out, tables, curve);
    return cudaGetLastError();
}

// =================================================================================================
// launchers
// =================================================================================================
static SampleArgs make_sample_args(const SampleSpec& sp) {
    SampleArgs a{};
    a.anims = sp.anims;
    a.write_untracked = sp.write_untracked ? 1u : 0u;
    a.range_mode = sp.range_mode ? 1u : 0u;
    a.frame_stride = sp.frame_stride;
    a.time_mode = sp.time_mode ? 1u : 0u;
    a.by_value = (sp.frame_by_value || sp.time_by_value) ? 1u : 0u;
    a.frame0 = sp.frame_by_value ? *sp.frame_by_value : 0u;
    a.time0 = sp.time_by_value ? *sp.time_by_value : 0.0;
    a.n_inline = (sp.frame_ids_inline && sp.n_inline <= kInlineFrameIds) ? sp.n_inline : 0u;
    for (uint32_t i = 0; i < a.n_inline; ++i) a.inline_ids[i] = sp.frame_ids_inline[i];
    return a;
}

cudaError_t launch_pose_sample(cudaStream_t st, const DevModel& M, const DevFrames& F, const SampleSpec& sp) {
    const uint32_t items = M.nb + M.nm;
    if (items == 0 || F.n_slots == 0) return cudaSuccess;
    dim3 grid((items + 127) / 128, F.n_slots);
    pose_sample_kernel<<<grid, 128, 0, st>>>(M, F, make_sample_args(sp));
    return cudaGetLastError();
}

// One-slot objects (an interactive Poser: sampling -> hierarchy -> skinning, each a few microseconds) launch the hierarchy
// and the skinning kernel with programmatic stream serialization: the kernel becomes resident and fetches the model's
// static data while its predecessor still runs, and passes its pdl_wait() the moment that one has finished.
// MMDGPU_PDL=0 switches it off.
static bool use_pdl(const DevFrames& F) {
    static const bool on = [] { const char* e = std::getenv("MMDGPU_PDL"); return !(e && e[0] == '0'); }();
    return on && F.n_slots == 1;
}
template <class... KArgs>
static cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, KArgs... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

cudaError_t launch_hierarchy(cudaStream_t st, const DevModel& M, const DevFrames& F, uint32_t wave_lo, uint32_t wave_hi,
                             bool prologue) {
    if (F.n_slots == 0) return cudaSuccess;
    const size_t cta_smem = hier_cta_smem_bytes(M.nb, M.n_link_slots, M.n_morph_slots, M.n_ops, M.n_waves);
    static const bool force_global = std::getenv("MMDGPU_FORCE_FALLBACKS") != nullptr;  // test knob
    if (cta_smem <= kHierCtaSmemLimit && !force_global) {
        // small skeletons: narrower CTAs, so that more slots are resident per SM (a CCD IK solve is one thread)
        const uint32_t threads = M.nb <= 512 ? 128u : kHierCtaThreads;
        const bool pdl = use_pdl(F);
        if (M.ik_nested) return launch_kernel(hierarchy_cta_kernel<true>, dim3(F.n_slots), dim3(threads), cta_smem, st, pdl, M, F, wave_lo, wave_hi, prologue ? 1u : 0u);
        return launch_kernel(hierarchy_cta_kernel<false>, dim3(F.n_slots), dim3(threads), cta_smem, st, pdl, M, F, wave_lo, wave_hi, prologue ? 1u : 0u);
    }
    const uint32_t blocks = (F.n_slots + kHierWarps - 1) / kHierWarps;
    if (M.ik_nested) hierarchy_kernel<true><<<blocks, 32 * kHierWarps, 0, st>>>(M, F, wave_lo, wave_hi, prologue ? 1u : 0u);
    else hierarchy_kernel<false><<<blocks, 32 * kHierWarps, 0, st>>>(M, F, wave_lo, wave_hi, prologue ? 1u : 0u);
    return cudaGetLastError();
}

bool hierarchy_uses_cta_kernel(const DevModel& M) {
    static const bool force_global = std::getenv("MMDGPU_FORCE_FALLBACKS") != nullptr;
    return !force_global && hier_cta_smem_bytes(M.nb, M.n_link_slots, M.n_morph_slots, M.n_ops, M.n_waves) <= kHierCtaSmemLimit;
}

size_t hierarchy_flat_smem_bytes(const DevModel& M) { return (size_t)kFlatThreads * M.ik_img_max_region * sizeof(float4); }

cudaError_t launch_hierarchy_wave_flat(cudaStream_t st, const DevModel& M, const DevFrames& F, uint32_t wave, uint32_t n_ops) {
    const uint64_t threads = (uint64_t)n_ops * F.n_slots;
    if (threads == 0) return cudaSuccess;
    hierarchy_flat_kernel<<<(unsigned)((threads + kFlatThreads - 1) / kFlatThreads), kFlatThreads, hierarchy_flat_smem_bytes(M), st>>>(M, F, wave);
    return cudaGetLastError();
}

static bool skin_runs_pair_kernel(const DevModel& M) {
    static const bool scalar = [] { const char* e = std::getenv("MMDGPU_SKIN_SCALAR"); return e && e[0] == '1'; }();
    return !M.extensions && !M.global_palette && !scalar;
}
static size_t skin_smem_rest(const DevModel& M) {   // palettes and rates, both double-buffered
    return 2 * (size_t)kSlotGroup * skin_pal_bytes(M.max_tile_bones, M.extensions != 0) + 2 * (size_t)M.n_nodes_pad * 16;
}
// The packed-pair kernel stages the 32-byte records too (16 KB per slot, one bulk copy per tile and slot) as long as three
// CTAs still fit an SM (228 KB, 1 KB reserved per CTA); larger tile palettes / more morph nodes keep the direct 256-bit
// stores.  MMDGPU_SOKOL_STAGED=0|1 forces either form (tests, A/B).
bool skin_sokol_staged(const DevModel& M, const DevFrames& F) {
    static const int forced = [] { const char* e = std::getenv("MMDGPU_SOKOL_STAGED"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    if (!skin_runs_pair_kernel(M)) return false;
    // measured (profiles/r02_experiments.md): +6 % on a bake of the 1 M-vertex model, -14 % on the crowd of 50 k-vertex
    // instances, whose lighter per-vertex work leaves the extra selects, moves and barriers of the staged form exposed
    if (forced < 0 && F.n_frames <= 1) return false;
    const size_t bytes = (size_t)kSlotGroup * skin_pair_stage_bytes(MMDGPU_LAYOUT_INTERLEAVED_SOKOL32, true) + skin_smem_rest(M);
    if (forced >= 0) return forced == 1 && bytes + 1024 <= 227 * 1024;
    return 3 * (bytes + 1024) <= 228 * 1024;
}

// upper bound over the forms a frames object of this model may run (load-time check)
size_t skin_smem_bytes(const DevModel& M, int layout) {
    const uint32_t stage = skin_runs_pair_kernel(M) ? skin_pair_stage_bytes(layout, false) : skin_stage_bytes(layout, M.extensions != 0);
    return (size_t)kSlotGroup * stage + skin_smem_rest(M);
}
static size_t skin_launch_smem_bytes(const DevModel& M, const DevFrames& F, int layout) {
    if (!skin_runs_pair_kernel(M)) return skin_smem_bytes(M, layout);
    return (size_t)kSlotGroup * skin_pair_stage_bytes(layout, skin_sokol_staged(M, F)) + skin_smem_rest(M);
}

// The dynamic shared-memory opt-in is an attribute of the kernel function (per device), not of a launch: every model
// of the process shares it.  It is therefore raised to the device limit once, never to one model's requirement (a
// smaller model loaded later would otherwise lower it under an earlier, larger one).
template <int LAYOUT, bool EXT, bool PALG>
static cudaError_t skin_opt_in(int limit) {
    return cudaFuncSetAttribute(skin_kernel<LAYOUT, EXT, PALG>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
}

cudaError_t prepare_skin_kernels(const DevModel& M) {
    (void)M;
    int dev = 0, limit = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    const int hier = limit < (int)kHierCtaSmemLimit ? limit : (int)kHierCtaSmemLimit;
    if ((e = cudaFuncSetAttribute(hierarchy_cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hier)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(hierarchy_cta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hier)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(hierarchy_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hier)) != cudaSuccess) return e;
    constexpr int SOA = MMDGPU_LAYOUT_SOA_POS_NRM, I32 = MMDGPU_LAYOUT_INTERLEAVED_SOKOL32;
    if ((e = skin_opt_in<SOA, true, false>(limit)) != cudaSuccess) return e;
    if ((e = skin_opt_in<I32, true, false>(limit)) != cudaSuccess) return e;
    if ((e = skin_opt_in<SOA, false, true>(limit)) != cudaSuccess) return e;
    if ((e = skin_opt_in<I32, false, true>(limit)) != cudaSuccess) return e;
    if ((e = skin_opt_in<SOA, false, false>(limit)) != cudaSuccess) return e;
    if ((e = skin_opt_in<I32, false, false>(limit)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(skin_pair_kernel<SOA, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(skin_pair_kernel<I32, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(skin_pair_kernel<I32, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
}

cudaError_t launch_skin(cudaStream_t st, const DevModel& M, const DevFrames& F, int layout, uint32_t slots_per_cta) {
    if (F.n_slots == 0 || M.n_tiles == 0) return cudaSuccess;
    slots_per_cta = (std::max<uint32_t>(slots_per_cta, 1u) + kSlotGroup - 1) / kSlotGroup * kSlotGroup;  // whole slot groups
    const uint32_t n_chunks = (F.n_slots + slots_per_cta - 1) / slots_per_cta;
    const uint32_t grid = M.n_tiles * n_chunks;
    const size_t smem = skin_launch_smem_bytes(M, F, layout);
    const bool soa = layout == MMDGPU_LAYOUT_SOA_POS_NRM;
    constexpr int SOA = MMDGPU_LAYOUT_SOA_POS_NRM, I32 = MMDGPU_LAYOUT_INTERLEAVED_SOKOL32;
#define MMDGPU_LAUNCH_SKIN(EXT, PALG)                                                                                  \
    do {                                                                                                               \
        if (soa) skin_kernel<SOA, EXT, PALG><<<grid, kSkinThreads, smem, st>>>(M, F, slots_per_cta, n_chunks);         \
        else skin_kernel<I32, EXT, PALG><<<grid, kSkinThreads, smem, st>>>(M, F, slots_per_cta, n_chunks);             \
    } while (0)
    // MMDGPU_SKIN_SCALAR=1 (test / experiment knob): the scalar kernel instead of the packed-pair one
    static const bool scalar = [] { const char* e = std::getenv("MMDGPU_SKIN_SCALAR"); return e && e[0] == '1'; }();
    if (M.extensions) MMDGPU_LAUNCH_SKIN(true, false);
    else if (M.global_palette) MMDGPU_LAUNCH_SKIN(false, true);
    else if (scalar) MMDGPU_LAUNCH_SKIN(false, false);
    else {
        // (the scalar kernels above read the palette through the read-only path and are never launched programmatically)
        const bool pdl = use_pdl(F);
        if (soa) return launch_kernel(skin_pair_kernel<SOA, 3, true>, dim3(grid), dim3(kSkinThreads), smem, st, pdl, M, F, slots_per_cta, n_chunks, -0.0f, 1.0f);
        if (skin_sokol_staged(M, F)) return launch_kernel(skin_pair_kernel<I32, 3, true>, dim3(grid), dim3(kSkinThreads), smem, st, pdl, M, F, slots_per_cta, n_chunks, -0.0f, 1.0f);
        return launch_kernel(skin_pair_kernel<I32, 3, false>, dim3(grid), dim3(kSkinThreads), smem, st, pdl, M, F, slots_per_cta, n_chunks, -0.0f, 1.0f);
    }
#undef MMDGPU_LAUNCH_SKIN
    return cudaGetLastError();
}

}  // namespace mmdgpu
