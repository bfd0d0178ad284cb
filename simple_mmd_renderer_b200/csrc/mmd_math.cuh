// mmd_math.cuh — device restatement of the libmmd fp32 expression trees used on the deformation path.
//
// Every function keeps libmmd's association order (L/ = 3rd_party/libmmd/include/mmd/ of the reference) and is
// compiled with -fmad=false so that no multiply-add is contracted: FMA contraction changes results beyond
// the parity tolerance and CCD IK amplifies the difference.  Trigonometry and sqrt go
// through double exactly like libmmd's math:: wrappers (L/util/math.inl:27-45).
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace mmdgpu {
namespace dm {

#define MMD_DEV __device__ __forceinline__

constexpr float kEpsF = 1e-7f;   // T(mmd_math_const_eps)
constexpr double kEpsD = 1e-7;   // mmd_math_const_eps as written (double macro, L/util/math.inl:24)

// libmmd: (float)sqrt((double)x).  Rounding a double square root to float equals the correctly rounded float square
// root for every float (53 >= 2*24 + 2 bits; verified exhaustively, tools/micro/sqrt_identity.c), so the fp32
// instruction (IEEE, -prec-sqrt=true) is used.
MMD_DEV float m_sqrt(float x) { return sqrtf(x); }
// sin and cos of one argument share the range reduction; sincos() is bit-identical to separate sin() / cos() on this
// toolchain (verified over every float in [-64, 64] and a sweep of large arguments, tools/micro/sincos_check.cu).
MMD_DEV void m_sincos(float x, float& s, float& c) {
    double ds, dc;
    sincos((double)x, &ds, &dc);
    s = (float)ds;
    c = (float)dc;
}
MMD_DEV float m_sin(float x) { return (float)sin((double)x); }
MMD_DEV float m_cos(float x) { return (float)cos((double)x); }
MMD_DEV float m_asin(float x) { return (float)asin((double)x); }
MMD_DEV float m_acos(float x) { return (float)acos((double)x); }
MMD_DEV float m_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
// std::max(a,b) = (a<b)?b:a ; std::min(a,b) = (b<a)?b:a — NaN behaviour included
MMD_DEV float s_max(float a, float b) { return (a < b) ? b : a; }
MMD_DEV float s_min(float a, float b) { return (b < a) ? b : a; }
MMD_DEV float m_clamp(float x, float lo, float hi) { return s_min(s_max(x, lo), hi); }

struct Quat { float i, j, k, e; };   // memory order (x, y, z, w), L/util/math.inl:259-265
struct Vec3 { float x, y, z; };
// Affine part of a Matrix4f: rows 0..3, columns 0..2 (column 3 is (0,0,0,1) for finite inputs).
struct Mat43 { float m[4][3]; };

MMD_DEV Quat q_identity() { return Quat{0.0f, 0.0f, 0.0f, 1.0f}; }
MMD_DEV Quat q_from(const float4& v) { return Quat{v.x, v.y, v.z, v.w}; }
MMD_DEV float4 q_to4(const Quat& q) { return make_float4(q.i, q.j, q.k, q.e); }

// Quaternion::operator*, L/util/math_impl.inl:510-517
MMD_DEV Quat q_mul(const Quat& a, const Quat& q) {
    Quat r;
    r.i = (a.e * q.i + a.i * q.e + a.j * q.k) - a.k * q.j;
    r.j = (a.e * q.j + a.j * q.e + a.k * q.i) - a.i * q.k;
    r.k = (a.e * q.k + a.i * q.j + a.k * q.e) - a.j * q.i;
    r.e = a.e * q.e - (a.i * q.i + a.j * q.j + a.k * q.k);
    return r;
}
// Quaternion::Inverse, L/util/math_impl.inl:474-477
MMD_DEV Quat q_inverse(const Quat& q) {
    float n = 1.0f / (q.i * q.i + q.j * q.j + q.k * q.k + q.e * q.e);
    return Quat{(-q.i) * n, (-q.j) * n, (-q.k) * n, q.e * n};
}
// Quaternion SLerp specialisation, L/util/math_impl.inl:1312-1340
MMD_DEV Quat q_slerp(const Quat& a, const Quat& b, float l) {
    float comega = a.e * b.e + a.i * b.i + a.j * b.j + a.k * b.k;
    bool flip = comega < 0.0f;
    if (flip) comega = -comega;
    float omega = m_acos(comega);
    if (omega > kEpsF) {
        float rs = 1.0f / m_sin(omega);
        float p = m_sin((1.0f - l) * omega) * rs;
        l = m_sin(l * omega) * rs;
        if (flip) l = -l;
        return Quat{a.i * p + b.i * l, a.j * p + b.j * l, a.k * p + b.k * l, a.e * p + b.e * l};
    }
    return a;
}
// NLerpProxy<Vector4f>::operator[], L/util/math_impl.inl:1265-1277 (key-frame rotations, motion_impl.inl:312)
MMD_DEV float4 v4_nlerp(const float4& a, const float4& b, float l) {
    if (l < kEpsF) return a;
    if (l > (1.0f - kEpsF)) return b;
    const float dot = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    float4 v;
    if (dot < 0.0f) {
        v.x = (1.0f - l) * a.x - l * b.x; v.y = (1.0f - l) * a.y - l * b.y;
        v.z = (1.0f - l) * a.z - l * b.z; v.w = (1.0f - l) * a.w - l * b.w;
    } else {
        v.x = (1.0f - l) * a.x + l * b.x; v.y = (1.0f - l) * a.y + l * b.y;
        v.z = (1.0f - l) * a.z + l * b.z; v.w = (1.0f - l) * a.w + l * b.w;
    }
    const float nn = 1.0f / m_sqrt(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    return make_float4(v.x * nn, v.y * nn, v.z * nn, v.w * nn);
}
// Quaternion::ToRotateMatrix, L/util/math_impl.inl:540-563 (rows 0..2; row 3 is set by the caller)
MMD_DEV void q_to_rows(const Quat& q, Mat43& M) {
    float ii = q.i * q.i, jj = q.j * q.j, kk = q.k * q.k;
    float ij = q.i * q.j, jk = q.j * q.k, ki = q.i * q.k;
    float ie = q.i * q.e, je = q.j * q.e, ke = q.k * q.e;
    M.m[0][0] = 1.0f - 2.0f * (jj + kk); M.m[0][1] = 2.0f * (ij + ke); M.m[0][2] = 2.0f * (ki - je);
    M.m[1][0] = 2.0f * (ij - ke); M.m[1][1] = 1.0f - 2.0f * (kk + ii); M.m[1][2] = 2.0f * (jk + ie);
    M.m[2][0] = 2.0f * (ki + je); M.m[2][1] = 2.0f * (jk - ie); M.m[2][2] = 1.0f - 2.0f * (ii + jj);
}
// Matrix4x4::operator*, L/util/math_impl.inl:984-1003, restricted to the 12 affine elements.  The fourth
// term of every sum is kept (a[r][3] is +0 for r < 3 and 1 for r = 3) so that signed zeros come out as in
// libmmd's full 4x4 product.
MMD_DEV Mat43 m_mul(const Mat43& a, const Mat43& b) {
    Mat43 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a3 = (i == 3) ? 1.0f : 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            r.m[i][c] = a.m[i][0] * b.m[0][c] + a.m[i][1] * b.m[1][c] + a.m[i][2] * b.m[2][c] + a3 * b.m[3][c];
    }
    return r;
}
MMD_DEV Mat43 m_identity() {
    Mat43 r;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) r.m[i][c] = (i == c) ? 1.0f : 0.0f;
    return r;
}
// Vector3D::Normalize, L/util/math_impl.inl:393-400
MMD_DEV Vec3 v_normalize(const Vec3& v) {
    float n = 1.0f / m_sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return Vec3{v.x * n, v.y * n, v.z * n};
}
MMD_DEV float v_dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// AxisToQuaternion, L/util/math_impl.inl:1047-1058
MMD_DEV Quat axis_to_quat(const Vec3& axis, float angle) {
    float norm = m_sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
    if (norm < kEpsF) return q_identity();
    angle *= 0.5f;
    float sn, cs;
    m_sincos(angle, sn, cs);
    float s = sn / norm;
    return Quat{s * axis.x, s * axis.y, s * axis.z, cs};
}
// QuaternionTo{ZXY,XYZ,YZX}, L/util/math_impl.inl:1123-1137, 1059-1073, 1107-1121.  order: 0 YZX 1 ZXY 2 XYZ
MMD_DEV Vec3 quat_to_euler(int order, const Quat& q) {
    float ii = q.i * q.i, jj = q.j * q.j, kk = q.k * q.k;
    float ei = q.e * q.i, ej = q.e * q.j, ek = q.e * q.k;
    float ij = q.i * q.j, ik = q.i * q.k, jk = q.j * q.k;
    Vec3 r;
    if (order == 1) {
        r.x = m_asin(2.0f * (ei + jk));
        r.y = m_atan2(2.0f * (ej - ik), 1.0f - 2.0f * (ii + jj));
        r.z = m_atan2(2.0f * (ek - ij), 1.0f - 2.0f * (ii + kk));
    } else if (order == 2) {
        r.x = m_atan2(2.0f * (ei - jk), 1.0f - 2.0f * (ii + jj));
        r.y = m_asin(2.0f * (ej + ik));
        r.z = m_atan2(2.0f * (ek - ij), 1.0f - 2.0f * (jj + kk));
    } else {
        r.x = m_atan2(2.0f * (ei - jk), 1.0f - 2.0f * (ii + kk));
        r.y = m_atan2(2.0f * (ej - ik), 1.0f - 2.0f * (jj + kk));
        r.z = m_asin(2.0f * (ek + ij));
    }
    return r;
}
// {ZXY,XYZ,YZX}ToQuaternion, L/util/math_impl.inl:1212-1224, 1156-1168, 1198-1210
MMD_DEV Quat euler_to_quat(int order, const Vec3& eu) {
    float cx, sx, cy, sy, cz, sz;
    m_sincos(eu.x * 0.5f, sx, cx);
    m_sincos(eu.y * 0.5f, sy, cy);
    m_sincos(eu.z * 0.5f, sz, cz);
    Quat q;
    if (order == 1) {
        q.e = cx * cy * cz - sx * sy * sz;
        q.i = sx * cy * cz - cx * sy * sz;
        q.j = cx * sy * cz + sx * cy * sz;
        q.k = cx * cy * sz + sx * sy * cz;
    } else if (order == 2) {
        q.e = cx * cy * cz - sx * sy * sz;
        q.i = sx * cy * cz + cx * sy * sz;
        q.j = cx * sy * cz - sx * cy * sz;
        q.k = sx * sy * cz + cx * cy * sz;
    } else {
        q.e = cx * cy * cz - sx * sy * sz;
        q.i = sx * cy * cz + cx * sy * sz;
        q.j = cx * sy * cz + sx * cy * sz;
        q.k = cx * cy * sz - sx * sy * cz;
    }
    return q;
}
// LimitEulerAngle, L/motion/poser_impl.inl:178-193
MMD_DEV float limit_one(float v, float lo, float hi, bool ikt) {
    if (v < lo) {
        float tf = 2 * lo - v;
        v = (tf <= hi && ikt) ? tf : lo;
    }
    if (v > hi) {
        float tf = 2 * hi - v;
        v = (tf >= lo && ikt) ? tf : hi;
    }
    return v;
}
// Bezier::operator[], L/util/math_impl.inl:1372-1384
MMD_DEV float bezier_at(const float* __restrict__ tables, uint32_t curve, float x) {
    if (curve == 0xFFFFFFFFu) return x;
    const float* tab = tables + (size_t)curve * 32;
    x *= 31.0f;
    unsigned long long ix = (unsigned long long)x;
    float r = x - (float)ix;
    if (ix < 31ull) return (1.0f - r) * tab[ix] + r * tab[ix + 1];
    return tab[31];
}

}  // namespace dm
}  // namespace mmdgpu
