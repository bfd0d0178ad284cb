/*
 * TEST INFRASTRUCTURE — not product code.  Nothing under simple_mmd_renderer_b200/ links, imports or
 * executes this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
 *
 * mmd_oracle.c — plain-C restatement of libmmd's per-frame deformation path on flat arrays.
 *
 * Every function cites the libmmd source it follows (paths relative to
 * /root/reference/3rd_party/libmmd/include/mmd/, abbreviated L/).  It is a sequential, literal
 * restatement (no wave schedule, no CSR): the point is an independent check of the device code.
 * Pinned: tests/test_oracle_pin.py compares it bit-for-bit with libmmd itself
 * (oracle/_ref/libmmd_ref.so) wherever /root/reference is mounted, and with the committed fixtures
 * under tests/golden/ (generated from libmmd by tests/golden/make_golden.py) everywhere.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (FMA contraction changes the result).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/mmdgpu.h"

#define EPS_D 1e-7                 /* mmd_math_const_eps, a double macro (L/util/math.inl:24) */
#define EPS_F ((float)1e-7)
#define PI_D 3.141592653589793238462643383279502884

#define EXPORT __attribute__((visibility("default")))

/* ---- math:: wrappers, L/util/math.inl:27-45: double libm, result rounded to float ---------------- */
static float m_sqrt(float x) { return (float)sqrt((double)x); }
static float m_sin(float x) { return (float)sin((double)x); }
static float m_cos(float x) { return (float)cos((double)x); }
static float m_asin(float x) { return (float)asin((double)x); }
static float m_acos(float x) { return (float)acos((double)x); }
static float m_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
/* std::max(a,b) = (a<b)?b:a ; std::min(a,b) = (b<a)?b:a */
static float s_max(float a, float b) { return (a < b) ? b : a; }
static float s_min(float a, float b) { return (b < a) ? b : a; }
static float m_clamp(float x, float lo, float hi) { return s_min(s_max(x, lo), hi); }

/* ---- Quaternion (x,y,z,w) = (i,j,k,e) ------------------------------------------------------------ */
/* L/util/math_impl.inl:510-517 */
static void q_mul(const float* a, const float* b, float* o) {
    float i = a[0], j = a[1], k = a[2], e = a[3];
    float qi = b[0], qj = b[1], qk = b[2], qe = b[3];
    float r0 = (e * qi + i * qe + j * qk) - k * qj;
    float r1 = (e * qj + j * qe + k * qi) - i * qk;
    float r2 = (e * qk + i * qj + k * qe) - j * qi;
    float r3 = e * qe - (i * qi + j * qj + k * qk);
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
}
/* L/util/math_impl.inl:474-477 */
static void q_inverse(const float* q, float* o) {
    float n = 1.0f / (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    o[0] = (-q[0]) * n; o[1] = (-q[1]) * n; o[2] = (-q[2]) * n; o[3] = q[3] * n;
}
/* L/util/math_impl.inl:540-563 — full 4x4, row-major */
static void q_to_matrix(const float* q, float* m) {
    float i = q[0], j = q[1], k = q[2], e = q[3];
    float ii = i * i, jj = j * j, kk = k * k, ij = i * j, jk = j * k, ki = i * k, ie = i * e, je = j * e, ke = k * e;
    m[0] = 1.0f - 2.0f * (jj + kk); m[1] = 2.0f * (ij + ke); m[2] = 2.0f * (ki - je);
    m[4] = 2.0f * (ij - ke); m[5] = 1.0f - 2.0f * (kk + ii); m[6] = 2.0f * (jk + ie);
    m[8] = 2.0f * (ki + je); m[9] = 2.0f * (jk - ie); m[10] = 1.0f - 2.0f * (ii + jj);
    m[3] = m[7] = m[11] = m[12] = m[13] = m[14] = 0.0f;
    m[15] = 1.0f;
}
/* L/util/math_impl.inl:984-1003 — every element a*b+c*d+e*f+g*h left to right */
static void m_mul(const float* a, const float* b, float* o) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j] + a[4 * i + 3] * b[12 + j];
    memcpy(o, r, sizeof r);
}
static void m_identity(float* m) {
    memset(m, 0, 16 * sizeof(float));
    m[0] = m[5] = m[10] = m[15] = 1.0f;
}
/* Quaternion SLerp specialisation, L/util/math_impl.inl:1312-1340 */
static void q_slerp(const float* a, const float* b, float l, float* o) {
    float comega = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    int flip = comega < 0.0f;
    if (flip) comega = -comega;
    float omega = m_acos(comega);
    if (omega > EPS_F) {
        float rs = 1.0f / m_sin(omega);
        float p = m_sin((1.0f - l) * omega) * rs;
        l = m_sin(l * omega) * rs;
        if (flip) l = -l;
        for (int c = 0; c < 4; ++c) o[c] = a[c] * p + b[c] * l;
    } else {
        for (int c = 0; c < 4; ++c) o[c] = a[c];
    }
}
/* Vector3D::Normalize, L/util/math_impl.inl:393-400 */
static void v3_normalize(const float* v, float* o) {
    float n = 1.0f / m_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    o[0] = v[0] * n; o[1] = v[1] * n; o[2] = v[2] * n;
}
static float v3_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* AxisToQuaternion, L/util/math_impl.inl:1047-1058 */
static void axis_to_quat(const float* axis, float angle, float* o) {
    float norm = m_sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    if (norm < EPS_F) {
        o[0] = o[1] = o[2] = 0.0f; o[3] = 1.0f;
    } else {
        angle *= 0.5f;
        float s = m_sin(angle) / norm;
        o[0] = s * axis[0]; o[1] = s * axis[1]; o[2] = s * axis[2];
        o[3] = m_cos(angle);
    }
}
/* Euler conversions, L/util/math_impl.inl:1059-1073, 1107-1137, 1156-1168, 1198-1224.  order: 0 YZX 1 ZXY 2 XYZ */
static void quat_to_euler(int order, const float* q, float* r) {
    float i = q[0], j = q[1], k = q[2], e = q[3];
    float ii = i * i, jj = j * j, kk = k * k, ei = e * i, ej = e * j, ek = e * k, ij = i * j, ik = i * k, jk = j * k;
    if (order == 1) { /* ZXY */
        r[0] = m_asin(2.0f * (ei + jk));
        r[1] = m_atan2(2.0f * (ej - ik), 1.0f - 2.0f * (ii + jj));
        r[2] = m_atan2(2.0f * (ek - ij), 1.0f - 2.0f * (ii + kk));
    } else if (order == 2) { /* XYZ */
        r[0] = m_atan2(2.0f * (ei - jk), 1.0f - 2.0f * (ii + jj));
        r[1] = m_asin(2.0f * (ej + ik));
        r[2] = m_atan2(2.0f * (ek - ij), 1.0f - 2.0f * (jj + kk));
    } else { /* YZX */
        r[0] = m_atan2(2.0f * (ei - jk), 1.0f - 2.0f * (ii + kk));
        r[1] = m_atan2(2.0f * (ej - ik), 1.0f - 2.0f * (jj + kk));
        r[2] = m_asin(2.0f * (ek + ij));
    }
}
static void euler_to_quat(int order, const float* eu, float* q) {
    float cx = m_cos(eu[0] * 0.5f), sx = m_sin(eu[0] * 0.5f);
    float cy = m_cos(eu[1] * 0.5f), sy = m_sin(eu[1] * 0.5f);
    float cz = m_cos(eu[2] * 0.5f), sz = m_sin(eu[2] * 0.5f);
    if (order == 1) { /* ZXY */
        q[3] = cx * cy * cz - sx * sy * sz;
        q[0] = sx * cy * cz - cx * sy * sz;
        q[1] = cx * sy * cz + sx * cy * sz;
        q[2] = cx * cy * sz + sx * sy * cz;
    } else if (order == 2) { /* XYZ */
        q[3] = cx * cy * cz - sx * sy * sz;
        q[0] = sx * cy * cz + cx * sy * sz;
        q[1] = cx * sy * cz - sx * cy * sz;
        q[2] = sx * sy * cz + cx * cy * sz;
    } else { /* YZX */
        q[3] = cx * cy * cz - sx * sy * sz;
        q[0] = sx * cy * cz + cx * sy * sz;
        q[1] = cx * sy * cz + sx * cy * sz;
        q[2] = cx * cy * sz - sx * sy * cz;
    }
}
/* LimitEulerAngle, L/motion/poser_impl.inl:178-193 */
static void limit_euler(float* r, const float* lo, const float* hi, int ikt) {
    for (int i = 0; i < 3; ++i) {
        if (r[i] < lo[i]) {
            float tf = 2 * lo[i] - r[i];
            if (tf <= hi[i] && ikt) r[i] = tf; else r[i] = lo[i];
        }
        if (r[i] > hi[i]) {
            float tf = 2 * hi[i] - r[i];
            if (tf >= lo[i] && ikt) r[i] = tf; else r[i] = hi[i];
        }
    }
}

/* ---- Bezier, L/util/math_impl.inl:1350-1428 ------------------------------------------------------ */
typedef struct { int linear; float tab[32]; } bezier;
static float bezier_interpolate(float c0x, float c0y, float c1x, float c1y, float x) {
    float l = 0.0f, r = 1.0f, m, lm = 0.0f, rm;
    for (int i = 0; i < 32; ++i) {
        lm = (l + r) * 0.5f;
        rm = 1.0f - lm;
        m = lm * (rm * (rm * c0x + lm * c1x) + lm * lm);
        if (fabsf(m - x) < EPS_F) break;
        if (m > x) r = lm; else l = lm;
    }
    rm = 1.0f - lm;
    return lm * (rm * (rm * c0y + lm * c1y) + lm * lm);
}
/* control bytes -> SetC -> presample; L/reader/vmd_reader_impl.inl:29-37, math_impl.inl:1393-1408 */
static void bezier_set(bezier* b, const int8_t c[4]) {
    const float r = 1.0f / 127.0f;
    float c0x = (c[0] * r) * 3.0f, c0y = (c[1] * r) * 3.0f, c1x = (c[2] * r) * 3.0f, c1y = (c[3] * r) * 3.0f;
    if (c0x == c0y && c1x == c1y) {
        b->linear = 1;
    } else {
        b->linear = 0;
        for (int i = 0; i < 32; ++i) b->tab[i] = bezier_interpolate(c0x, c0y, c1x, c1y, (float)i / 31.0f);
    }
}
/* Bezier::operator[], L/util/math_impl.inl:1372-1384 */
static float bezier_at(const bezier* b, float x) {
    if (b->linear) return x;
    x *= 31.0f;
    size_t ix = (size_t)x;
    float r = x - (float)ix;
    if (ix < 31) return (1.0f - r) * b->tab[ix] + r * b->tab[ix + 1];
    return b->tab[31];
}

/* ---- session ---------------------------------------------------------------------------------- */
typedef struct {
    uint32_t frame;
    float T[3], R[4];
    bezier ip[4];
} bkey;
typedef struct { uint32_t frame; float w; } mkey;

typedef struct {
    /* model (copied) */
    uint32_t nv, nb, nm, nlinks;
    float *pos, *nrm, *uv;
    uint8_t* stype;     /* after Normalize: 0 BDEF1 1 BDEF2 2 BDEF4 3 SDEF */
    int32_t* sid;       /* 4nv */
    float* sw;          /* 4nv */
    float *bpos;
    int32_t *parent, *level;
    uint16_t* flags;
    /* Poser::Poser precomputation, L/motion/poser_impl.inl:30-105 */
    uint8_t *has_parent, *has_append, *app_rot, *app_trans, *has_ik, *is_link;
    int32_t* app_parent;
    float* app_ratio;
    float* local_offset;
    int32_t* ik_target;
    uint32_t *ik_iters, *ik_lbegin, *ik_lcount;
    float* ik_angle;
    int32_t* l_bone;
    uint8_t *l_limited, *l_fix, *l_order;
    float *l_min, *l_max;
    uint32_t n_pre, n_post;
    int32_t *order_pre, *order_post;
    /* morphs */
    uint8_t* mtype;
    uint32_t *mbegin, *mcount;
    mmdgpu_vertex_morph_entry* ve;
    mmdgpu_bone_morph_entry* be;
    mmdgpu_group_morph_entry* ge;
    /* motion */
    int has_motion;
    uint32_t n_btracks, n_mtracks;
    int32_t *bt_bone, *mt_morph;
    uint32_t *bt_begin, *bt_count, *mt_begin, *mt_count;
    bkey* bkeys;
    mkey* mkeys;
} model_t;

typedef struct {
    const model_t* m;
    float *R, *T, *rate;                        /* rotation_, translation_, morph_rates_ */
    float *morphR, *morphT, *totR, *totT, *preIK, *ikR;
    float *local, *skin;                        /* 16 floats per bone */
    float *vimg, *opos, *onrm;
} state_t;

struct port_session { model_t m; state_t s; };

static void* dup_mem(const void* p, size_t n) {
    void* r = malloc(n ? n : 1);
    if (p && n) memcpy(r, p, n); else if (n) memset(r, 0, n);
    return r;
}

static int cmp_order(const void* a, const void* b, void* ctx) {
    const model_t* m = (const model_t*)ctx;
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    uint32_t lx = (uint32_t)m->level[x], ly = (uint32_t)m->level[y];
    if (lx < ly) return -1;
    if (lx > ly) return 1;
    return (x < y) ? -1 : (x > y);
}
/* tiny insertion sort keeps this file free of non-standard qsort_r */
static void sort_order(const model_t* m, int32_t* v, uint32_t n) {
    for (uint32_t i = 1; i < n; ++i) {
        int32_t x = v[i];
        uint32_t j = i;
        while (j > 0 && cmp_order(&x, &v[j - 1], (void*)m) < 0) { v[j] = v[j - 1]; --j; }
        v[j] = x;
    }
}

/* Model::Normalize, L/model/model_impl.inl:406-452 */
static void normalize_skinning(model_t* m, const mmdgpu_model_desc* d) {
    for (uint32_t i = 0; i < m->nv; ++i) {
        int32_t* id = m->sid + 4 * i;
        float* w = m->sw + 4 * i;
        uint8_t t = d->skin_type[i];
        for (int k = 0; k < 4; ++k) { id[k] = d->bone_id[4 * i + k]; w[k] = d->weight[4 * i + k]; }
        if (t == MMDGPU_SKIN_QDEF || t > MMDGPU_SKIN_QDEF) t = MMDGPU_SKIN_BDEF4; /* libmmd has no QDEF */
        if (t == MMDGPU_SKIN_BDEF2) {
            if (w[0] == 0.0f) { id[0] = id[1]; t = MMDGPU_SKIN_BDEF1; }
            else if (w[0] == 1.0f) { t = MMDGPU_SKIN_BDEF1; }
        } else if (t == MMDGPU_SKIN_SDEF) {
            int32_t b0 = id[0], b1 = id[1];
            if (m->parent[b0] != b1 && m->parent[b1] != b0) {
                if (w[0] == 0.0f) { id[0] = id[1]; t = MMDGPU_SKIN_BDEF1; }
                else if (w[0] == 1.0f) { t = MMDGPU_SKIN_BDEF1; }
                else t = MMDGPU_SKIN_BDEF2;
            }
        }
        m->stype[i] = t;
    }
}

static void build_model(model_t* m, const mmdgpu_model_desc* d) {
    uint32_t nv = d->n_vertices, nb = d->n_bones, nm = d->n_morphs, nl = d->n_ik_links;
    m->nv = nv; m->nb = nb; m->nm = nm; m->nlinks = nl;
    m->pos = dup_mem(d->position, 12u * nv);
    m->nrm = dup_mem(d->normal, 12u * nv);
    m->uv = dup_mem(d->uv, 8u * nv);
    m->stype = dup_mem(NULL, nv);
    m->sid = dup_mem(NULL, 16u * nv);
    m->sw = dup_mem(NULL, 16u * nv);
    m->bpos = dup_mem(d->bone_position, 12u * nb);
    m->parent = dup_mem(d->bone_parent, 4u * nb);
    m->level = dup_mem(d->bone_transform_level, 4u * nb);
    m->flags = dup_mem(d->bone_flags, 2u * nb);
    m->has_parent = dup_mem(NULL, nb); m->has_append = dup_mem(NULL, nb); m->app_rot = dup_mem(NULL, nb);
    m->app_trans = dup_mem(NULL, nb); m->has_ik = dup_mem(NULL, nb); m->is_link = dup_mem(NULL, nb);
    m->app_parent = dup_mem(NULL, 4u * nb); m->app_ratio = dup_mem(NULL, 4u * nb);
    m->local_offset = dup_mem(NULL, 12u * nb);
    m->ik_target = dup_mem(NULL, 4u * nb); m->ik_iters = dup_mem(NULL, 4u * nb);
    m->ik_lbegin = dup_mem(NULL, 4u * nb); m->ik_lcount = dup_mem(NULL, 4u * nb); m->ik_angle = dup_mem(NULL, 4u * nb);
    m->l_bone = dup_mem(d->ik_link_bone, 4u * nl);
    m->l_limited = dup_mem(d->ik_link_has_limit, nl);
    m->l_fix = dup_mem(NULL, nl); m->l_order = dup_mem(NULL, nl);
    m->l_min = dup_mem(NULL, 12u * nl); m->l_max = dup_mem(NULL, 12u * nl);
    m->order_pre = dup_mem(NULL, 4u * nb); m->order_post = dup_mem(NULL, 4u * nb);
    /* parent index: -1 (nil) or >= nb means "no parent" (poser_impl.inl:39-46); keep -1 for Normalize's compare */
    for (uint32_t b = 0; b < nb; ++b) {
        int32_t p = m->parent[b];
        if (p >= 0 && (uint32_t)p < nb) {
            m->has_parent[b] = 1;
            for (int k = 0; k < 3; ++k) m->local_offset[3 * b + k] = m->bpos[3 * b + k] - m->bpos[3 * p + k];
        } else {
            for (int k = 0; k < 3; ++k) m->local_offset[3 * b + k] = m->bpos[3 * b + k];
        }
        uint16_t fl = m->flags[b];
        m->app_rot[b] = (fl & MMDGPU_BONE_APPEND_ROTATE) != 0;
        m->app_trans[b] = (fl & MMDGPU_BONE_APPEND_TRANSLATE) != 0;
        if (m->app_rot[b] || m->app_trans[b]) {
            int32_t ap = d->bone_append_parent ? d->bone_append_parent[b] : -1;
            m->app_parent[b] = ap;
            if (ap >= 0 && (uint32_t)ap < nb) {
                m->has_append[b] = 1;
                m->app_ratio[b] = d->bone_append_ratio ? d->bone_append_ratio[b] : 0.0f;
            }
        }
        m->has_ik[b] = (fl & MMDGPU_BONE_HAS_IK) != 0;
        if (m->has_ik[b]) {
            m->ik_lbegin[b] = d->ik_link_begin[b];
            m->ik_lcount[b] = d->ik_link_count[b];
            for (uint32_t j = 0; j < m->ik_lcount[b]; ++j) {
                uint32_t l = m->ik_lbegin[b] + j;
                m->l_order[l] = 0; /* ORDER_YZX default, poser_impl.inl:64 */
                m->l_fix[l] = 0;   /* FIX_NONE */
                if (m->l_limited[l]) {
                    float* mn = m->l_min + 3 * l; float* mx = m->l_max + 3 * l;
                    for (int k = 0; k < 3; ++k) {
                        float lo = d->ik_link_lo[3 * l + k], hi = d->ik_link_hi[3 * l + k];
                        mn[k] = s_min(lo, hi); mx[k] = s_max(lo, hi);
                    }
                    /* poser_impl.inl:78-82: compared as double against -pi*0.5f */
                    if ((double)mn[0] > -PI_D * 0.5f && (double)mx[0] < PI_D * 0.5f) m->l_order[l] = 1;      /* ZXY */
                    else if ((double)mn[1] > -PI_D * 0.5f && (double)mx[1] < PI_D * 0.5f) m->l_order[l] = 2; /* XYZ */
                    /* poser_impl.inl:83-91: float abs, double compare */
                    int zx = (double)fabsf(mn[0]) < EPS_D && (double)fabsf(mx[0]) < EPS_D;
                    int zy = (double)fabsf(mn[1]) < EPS_D && (double)fabsf(mx[1]) < EPS_D;
                    int zz = (double)fabsf(mn[2]) < EPS_D && (double)fabsf(mx[2]) < EPS_D;
                    if (zx && zy && zz) m->l_fix[l] = 4;
                    else if (zy && zz) m->l_fix[l] = 1;
                    else if (zx && zz) m->l_fix[l] = 2;
                    else if (zx && zy) m->l_fix[l] = 3;
                }
                m->is_link[m->l_bone[l]] = 1;
            }
            m->ik_angle[b] = d->ik_angle_limit[b];
            uint32_t it = (uint32_t)d->ik_iterations[b];
            m->ik_iters[b] = it < 256u ? it : 256u;
            m->ik_target[b] = d->ik_target[b];
        }
        if (fl & MMDGPU_BONE_POST_PHYSICS) m->order_post[m->n_post++] = (int32_t)b;
        else m->order_pre[m->n_pre++] = (int32_t)b;
    }
    sort_order(m, m->order_pre, m->n_pre);
    sort_order(m, m->order_post, m->n_post);
    normalize_skinning(m, d);
    m->mtype = dup_mem(d->morph_type, nm);
    m->mbegin = dup_mem(d->morph_entry_begin, 4u * nm);
    m->mcount = dup_mem(d->morph_entry_count, 4u * nm);
    m->ve = dup_mem(d->vertex_morph_entries, sizeof(mmdgpu_vertex_morph_entry) * (size_t)d->n_vertex_morph_entries);
    m->be = dup_mem(d->bone_morph_entries, sizeof(mmdgpu_bone_morph_entry) * (size_t)d->n_bone_morph_entries);
    m->ge = dup_mem(d->group_morph_entries, sizeof(mmdgpu_group_morph_entry) * (size_t)d->n_group_morph_entries);
}

/* std::map<size_t, Keyframe> semantics: sorted by frame, equal frames keep the last record */
static uint32_t sort_dedupe_idx(const uint32_t* frames, uint32_t n, uint32_t* idx) {
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    for (uint32_t i = 1; i < n; ++i) { /* stable insertion sort by frame */
        uint32_t x = idx[i], j = i;
        while (j > 0 && frames[idx[j - 1]] > frames[x]) { idx[j] = idx[j - 1]; --j; }
        idx[j] = x;
    }
    uint32_t o = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (i + 1 < n && frames[idx[i + 1]] == frames[idx[i]]) continue; /* a later duplicate wins */
        idx[o++] = idx[i];
    }
    return o;
}

static void build_motion(model_t* m, const mmdgpu_anim_desc* a) {
    m->has_motion = 1;
    m->n_btracks = a->n_bone_tracks; m->n_mtracks = a->n_morph_tracks;
    m->bt_bone = dup_mem(a->bone_track_bone, 4u * a->n_bone_tracks);
    m->bt_begin = dup_mem(NULL, 4u * a->n_bone_tracks);
    m->bt_count = dup_mem(NULL, 4u * a->n_bone_tracks);
    m->mt_morph = dup_mem(a->morph_track_morph, 4u * a->n_morph_tracks);
    m->mt_begin = dup_mem(NULL, 4u * a->n_morph_tracks);
    m->mt_count = dup_mem(NULL, 4u * a->n_morph_tracks);
    m->bkeys = dup_mem(NULL, sizeof(bkey) * (size_t)a->n_bone_keys);
    m->mkeys = dup_mem(NULL, sizeof(mkey) * (size_t)a->n_morph_keys);
    uint32_t maxn = 1;
    for (uint32_t t = 0; t < a->n_bone_tracks; ++t) if (a->bone_track_key_count[t] > maxn) maxn = a->bone_track_key_count[t];
    for (uint32_t t = 0; t < a->n_morph_tracks; ++t) if (a->morph_track_key_count[t] > maxn) maxn = a->morph_track_key_count[t];
    uint32_t* idx = malloc(4u * maxn);
    uint32_t* fr = malloc(4u * maxn);
    uint32_t o = 0;
    for (uint32_t t = 0; t < a->n_bone_tracks; ++t) {
        uint32_t b0 = a->bone_track_key_begin[t], n = a->bone_track_key_count[t];
        for (uint32_t i = 0; i < n; ++i) fr[i] = a->bone_keys[b0 + i].frame;
        uint32_t k = sort_dedupe_idx(fr, n, idx);
        m->bt_begin[t] = o; m->bt_count[t] = k;
        for (uint32_t i = 0; i < k; ++i) {
            const mmdgpu_bone_key* s = &a->bone_keys[b0 + idx[i]];
            bkey* d = &m->bkeys[o++];
            d->frame = s->frame;
            memcpy(d->T, s->translation, 12); memcpy(d->R, s->rotation, 16);
            for (int c = 0; c < 4; ++c) bezier_set(&d->ip[c], s->interp[c]);
        }
    }
    o = 0;
    for (uint32_t t = 0; t < a->n_morph_tracks; ++t) {
        uint32_t b0 = a->morph_track_key_begin[t], n = a->morph_track_key_count[t];
        for (uint32_t i = 0; i < n; ++i) fr[i] = a->morph_keys[b0 + i].frame;
        uint32_t k = sort_dedupe_idx(fr, n, idx);
        m->mt_begin[t] = o; m->mt_count[t] = k;
        for (uint32_t i = 0; i < k; ++i) {
            m->mkeys[o].frame = a->morph_keys[b0 + idx[i]].frame;
            m->mkeys[o].w = a->morph_keys[b0 + idx[i]].weight;
            ++o;
        }
    }
    free(idx); free(fr);
}

static void state_init(state_t* s, const model_t* m) {
    s->m = m;
    uint32_t nb = m->nb, nv = m->nv;
    s->R = dup_mem(NULL, 16u * nb); s->T = dup_mem(NULL, 12u * nb); s->rate = dup_mem(NULL, 4u * m->nm);
    s->morphR = dup_mem(NULL, 16u * nb); s->morphT = dup_mem(NULL, 12u * nb);
    s->totR = dup_mem(NULL, 16u * nb); s->totT = dup_mem(NULL, 12u * nb);
    s->preIK = dup_mem(NULL, 16u * nb); s->ikR = dup_mem(NULL, 16u * nb);
    s->local = dup_mem(NULL, 64u * nb); s->skin = dup_mem(NULL, 64u * nb);
    s->vimg = dup_mem(NULL, 12u * nv); s->opos = dup_mem(NULL, 12u * nv); s->onrm = dup_mem(NULL, 12u * nv);
    for (uint32_t b = 0; b < nb; ++b) s->R[4 * b + 3] = 1.0f;
}
static void state_free(state_t* s) {
    free(s->R); free(s->T); free(s->rate); free(s->morphR); free(s->morphT); free(s->totR); free(s->totT);
    free(s->preIK); free(s->ikR); free(s->local); free(s->skin); free(s->vimg); free(s->opos); free(s->onrm);
}

/* ---- sampling: Motion::GetBonePose / GetMorphPose, L/motion/motion_impl.inl:255-319, 382-424 ---- */
static void sample_bone(const model_t* m, uint32_t t, uint32_t frame, float* T, float* R) {
    uint32_t n = m->bt_count[t];
    const bkey* k = m->bkeys + m->bt_begin[t];
    if (n == 0) { T[0] = T[1] = T[2] = 0; R[0] = R[1] = R[2] = 0; R[3] = 1; return; }
    const bkey* use = NULL;
    if (k[0].frame >= frame) use = &k[0];
    else if (k[n - 1].frame <= frame) use = &k[n - 1];
    if (!use) {
        uint32_t r = 0; /* upper_bound */
        while (k[r].frame <= frame) ++r;
        const bkey* rk = &k[r];
        const bkey* lk = &k[r - 1];
        if (lk->frame == frame) use = lk;
        else {
            float bary = (float)(frame - lk->frame) / (float)(rk->frame - lk->frame);
            for (int c = 0; c < 3; ++c) {
                float lam = bezier_at(&lk->ip[c], bary);
                T[c] = lk->T[c] * (1 - lam) + rk->T[c] * lam;
            }
            float l = bezier_at(&lk->ip[3], bary);
            /* NLerpProxy<Vector4f>, L/util/math_impl.inl:1265-1277 */
            if (l < EPS_F) { memcpy(R, lk->R, 16); return; }
            if (l > (1.0f - EPS_F)) { memcpy(R, rk->R, 16); return; }
            const float* a = lk->R; const float* b = rk->R;
            float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
            float v[4];
            if (dot < 0.0f) for (int c = 0; c < 4; ++c) v[c] = (1.0f - l) * a[c] - l * b[c];
            else for (int c = 0; c < 4; ++c) v[c] = (1.0f - l) * a[c] + l * b[c];
            float nn = 1.0f / m_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
            for (int c = 0; c < 4; ++c) R[c] = v[c] * nn;
            return;
        }
    }
    memcpy(T, use->T, 12); memcpy(R, use->R, 16);
}
static float sample_morph(const model_t* m, uint32_t t, uint32_t frame) {
    uint32_t n = m->mt_count[t];
    const mkey* k = m->mkeys + m->mt_begin[t];
    if (n == 0) return 0.0f;
    if (k[0].frame >= frame) return k[0].w;
    if (k[n - 1].frame <= frame) return k[n - 1].w;
    uint32_t r = 0;
    while (k[r].frame <= frame) ++r;
    if (k[r - 1].frame == frame) return k[r - 1].w;
    float bary = (float)(frame - k[r - 1].frame) / (float)(k[r].frame - k[r - 1].frame);
    float lam = bary; /* default-constructed Bezier is linear (math_impl.inl:1350-1354) */
    return k[r - 1].w * (1 - lam) + k[r].w * lam;
}

/* Motion::GetBonePose / GetMorphPose(name, double time), L/motion/motion_impl.inl:321-380, 426-470: the frame is
 * time*30 as a double, the bracket is upper_bound(size_t(dframe)), there is no "left key == frame" shortcut, and
 * the barycentre is computed in double and cast to float. */
static void sample_bone_time(const model_t* m, uint32_t t, double time, float* T, float* R) {
    uint32_t n = m->bt_count[t];
    const bkey* k = m->bkeys + m->bt_begin[t];
    if (n == 0) { T[0] = T[1] = T[2] = 0; R[0] = R[1] = R[2] = 0; R[3] = 1; return; }
    double dframe = time * 30.0;
    const bkey* use = NULL;
    if ((double)k[0].frame >= dframe) use = &k[0];
    else if ((double)k[n - 1].frame <= dframe) use = &k[n - 1];
    if (use) { memcpy(T, use->T, 12); memcpy(R, use->R, 16); return; }
    size_t key = (size_t)dframe;
    uint32_t r = 0;
    while (k[r].frame <= key) ++r;
    const bkey* rk = &k[r];
    const bkey* lk = &k[r - 1];
    float bary = (float)((dframe - (double)lk->frame) / (double)((size_t)rk->frame - (size_t)lk->frame));
    for (int c = 0; c < 3; ++c) {
        float lam = bezier_at(&lk->ip[c], bary);
        T[c] = lk->T[c] * (1 - lam) + rk->T[c] * lam;
    }
    float l = bezier_at(&lk->ip[3], bary);
    if (l < EPS_F) { memcpy(R, lk->R, 16); return; }
    if (l > (1.0f - EPS_F)) { memcpy(R, rk->R, 16); return; }
    const float* a = lk->R; const float* b = rk->R;
    float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    float v[4];
    if (dot < 0.0f) for (int c = 0; c < 4; ++c) v[c] = (1.0f - l) * a[c] - l * b[c];
    else for (int c = 0; c < 4; ++c) v[c] = (1.0f - l) * a[c] + l * b[c];
    float nn = 1.0f / m_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    for (int c = 0; c < 4; ++c) R[c] = v[c] * nn;
}
static float sample_morph_time(const model_t* m, uint32_t t, double time) {
    uint32_t n = m->mt_count[t];
    const mkey* k = m->mkeys + m->mt_begin[t];
    if (n == 0) return 0.0f;
    double dframe = time * 30.0;
    if ((double)k[0].frame >= dframe) return k[0].w;
    if ((double)k[n - 1].frame <= dframe) return k[n - 1].w;
    size_t key = (size_t)dframe;
    uint32_t r = 0;
    while (k[r].frame <= key) ++r;
    float bary = (float)((dframe - (double)k[r - 1].frame) / (double)((size_t)k[r].frame - (size_t)k[r - 1].frame));
    float lam = bary;
    return k[r - 1].w * (1 - lam) + k[r].w * lam;
}

/* ---- Poser -------------------------------------------------------------------------------------- */
/* Poser::ResetPosing's pose part, L/motion/poser_impl.inl:131-137 */
static void reset_poses(state_t* s) {
    const model_t* m = s->m;
    for (uint32_t i = 0; i < m->nm; ++i) s->rate[i] = 0;
    for (uint32_t b = 0; b < m->nb; ++b) {
        s->R[4 * b] = s->R[4 * b + 1] = s->R[4 * b + 2] = 0; s->R[4 * b + 3] = 1;
        s->T[3 * b] = s->T[3 * b + 1] = s->T[3 * b + 2] = 0;
    }
}
/* MotionPlayer::SeekFrame, L/motion/poser_impl.inl:539-546 */
static void seek_frame(state_t* s, uint32_t frame) {
    const model_t* m = s->m;
    if (!m->has_motion) return;
    for (uint32_t t = 0; t < m->n_mtracks; ++t) s->rate[m->mt_morph[t]] = sample_morph(m, t, frame);
    for (uint32_t t = 0; t < m->n_btracks; ++t) {
        int32_t b = m->bt_bone[t];
        sample_bone(m, t, frame, s->T + 3 * b, s->R + 4 * b);
    }
}

/* MotionPlayer::SeekTime, L/motion/poser_impl.inl:548-555 */
static void seek_time(state_t* s, double time) {
    const model_t* m = s->m;
    if (!m->has_motion) return;
    for (uint32_t t = 0; t < m->n_mtracks; ++t) s->rate[m->mt_morph[t]] = sample_morph_time(m, t, time);
    for (uint32_t t = 0; t < m->n_btracks; ++t) {
        int32_t b = m->bt_bone[t];
        sample_bone_time(m, t, time, s->T + 3 * b, s->R + 4 * b);
    }
}

static void set_local_from_rotation(state_t* s, uint32_t b) {
    const model_t* m = s->m;
    float* L = s->local + 16 * b;
    q_to_matrix(s->totR + 4 * b, L);
    for (int k = 0; k < 3; ++k) L[12 + k] = s->totT[3 * b + k] + m->local_offset[3 * b + k];
    if (m->has_parent[b]) m_mul(L, s->local + 16 * m->parent[b], L);
}

/* Poser::UpdateBoneTransform(size_t), L/motion/poser_impl.inl:142-311 */
static void update_bone(state_t* s, uint32_t b) {
    const model_t* m = s->m;
    float* totR = s->totR + 4 * b;
    float* totT = s->totT + 3 * b;
    q_mul(s->morphR + 4 * b, s->R + 4 * b, totR);
    for (int k = 0; k < 3; ++k) totT[k] = s->morphT[3 * b + k] + s->T[3 * b + k];
    if (m->has_append[b]) {
        int32_t ap = m->app_parent[b];
        if (m->app_rot[b]) {
            const float ident[4] = {0, 0, 0, 1};
            float sl[4];
            q_slerp(ident, s->totR + 4 * ap, m->app_ratio[b], sl);
            q_mul(totR, sl, totR);
        }
        if (m->app_trans[b])
            for (int k = 0; k < 3; ++k) totT[k] = totT[k] + m->app_ratio[b] * s->totT[3 * ap + k];
    }
    if (m->is_link[b]) {
        memcpy(s->preIK + 4 * b, totR, 16);
        q_mul(s->ikR + 4 * b, totR, totR);
    }
    set_local_from_rotation(s, b);

    if (!m->has_ik[b]) return;
    uint32_t nl = m->ik_lcount[b], lb = m->ik_lbegin[b];
    const int32_t* links = m->l_bone + lb;
    for (uint32_t i = 0; i < nl; ++i) { float* q = s->ikR + 4 * links[i]; q[0] = q[1] = q[2] = 0; q[3] = 1; }
    float ik_pos[3]; memcpy(ik_pos, s->local + 16 * b + 12, 12);
    for (uint32_t i = 0; i < nl; ++i) update_bone(s, (uint32_t)links[nl - i - 1]);
    int32_t tgt = m->ik_target[b];
    update_bone(s, (uint32_t)tgt);
    float tp[3]; memcpy(tp, s->local + 16 * tgt + 12, 12);
    float err[3] = {ik_pos[0] - tp[0], ik_pos[1] - tp[1], ik_pos[2] - tp[2]};
    if ((double)v3_dot(err, err) < EPS_D) return;
    uint32_t iters = m->ik_iters[b];
    uint32_t ikt = iters / 2;
    for (uint32_t i = 0; i < iters; ++i) {
        for (uint32_t j = 0; j < nl; ++j) {
            uint32_t l = lb + j;
            if (m->l_fix[l] == 4) continue;
            uint32_t lk = (uint32_t)links[j];
            const float* lp = s->local + 16 * lk + 12;
            float td[3] = {lp[0] - tp[0], lp[1] - tp[1], lp[2] - tp[2]};
            float id[3] = {lp[0] - ik_pos[0], lp[1] - ik_pos[1], lp[2] - ik_pos[2]};
            v3_normalize(td, td);
            v3_normalize(id, id);
            float ax[3]; /* Triple::operator*, math_impl.inl:260-266 */
            ax[0] = td[1] * id[2] - td[2] * id[1];
            ax[1] = td[2] * id[0] - td[0] * id[2];
            ax[2] = td[0] * id[1] - td[1] * id[0];
            for (int k = 0; k < 3; ++k) if ((double)fabsf(ax[k]) < EPS_D) ax[k] = (float)EPS_D;
            float P[16];
            if (m->has_parent[lk]) memcpy(P, s->local + 16 * m->parent[lk], 64); else m_identity(P);
            int fix = m->l_fix[l];
            if (m->l_limited[l] && fix != 0 && i < ikt) {
                int r = fix - 1; /* X->row0, Y->row1, Z->row2 */
                float d = v3_dot(ax, P + 4 * r);
                float sgn = (d >= 0.0f) ? 1.0f : -1.0f;
                ax[0] = ax[1] = ax[2] = 0.0f;
                ax[r] = sgn;
            } else {
                /* rotate(axis, P^T), math_impl.inl:1032-1038: component c = axis . P.row(c).xyz */
                float t0 = ax[0] * P[0] + ax[1] * P[1] + ax[2] * P[2];
                float t1 = ax[0] * P[4] + ax[1] * P[5] + ax[2] * P[6];
                float t2 = ax[0] * P[8] + ax[1] * P[9] + ax[2] * P[10];
                ax[0] = t0; ax[1] = t1; ax[2] = t2;
                v3_normalize(ax, ax);
            }
            float ang = s_min(m_acos(m_clamp(v3_dot(td, id), -1.0f, 1.0f)), m->ik_angle[b] * (float)(j + 1));
            float aq[4];
            axis_to_quat(ax, ang, aq);
            float* ikR = s->ikR + 4 * lk;
            q_mul(aq, ikR, ikR);
            if (m->l_limited[l]) {
                float lr[4], eu[3], inv[4];
                q_mul(ikR, s->preIK + 4 * lk, lr);
                quat_to_euler(m->l_order[l], lr, eu);
                limit_euler(eu, m->l_min + 3 * l, m->l_max + 3 * l, i < ikt);
                euler_to_quat(m->l_order[l], eu, lr);
                q_inverse(s->preIK + 4 * lk, inv);
                q_mul(lr, inv, ikR);
            }
            for (uint32_t k = 0; k <= j; ++k) {
                uint32_t c = (uint32_t)links[j - k];
                q_mul(s->ikR + 4 * c, s->preIK + 4 * c, s->totR + 4 * c);
                set_local_from_rotation(s, c);
            }
            update_bone(s, (uint32_t)tgt);
            memcpy(tp, s->local + 16 * tgt + 12, 12);
        }
        err[0] = ik_pos[0] - tp[0]; err[1] = ik_pos[1] - tp[1]; err[2] = ik_pos[2] - tp[2];
        if (v3_dot(err, err) < EPS_F) return;
    }
}

/* Poser::UpdateMorphTransform, L/motion/poser_impl.inl:328-360 */
static void update_morph(state_t* s, uint32_t idx, float rate) {
    const model_t* m = s->m;
    if ((double)rate < EPS_D) return;
    uint32_t b0 = m->mbegin[idx], n = m->mcount[idx];
    switch (m->mtype[idx]) {
    case MMDGPU_MORPH_GROUP:
        for (uint32_t i = 0; i < n; ++i) update_morph(s, m->ge[b0 + i].morph, m->ge[b0 + i].rate * rate);
        break;
    case MMDGPU_MORPH_VERTEX:
        for (uint32_t i = 0; i < n; ++i) {
            const mmdgpu_vertex_morph_entry* e = &m->ve[b0 + i];
            float* vi = s->vimg + 3 * (size_t)e->vertex;
            for (int k = 0; k < 3; ++k) vi[k] = vi[k] + e->offset[k] * rate;
        }
        break;
    case MMDGPU_MORPH_BONE:
        for (uint32_t i = 0; i < n; ++i) {
            const mmdgpu_bone_morph_entry* e = &m->be[b0 + i];
            float* mt = s->morphT + 3 * e->bone;
            float* mr = s->morphR + 4 * e->bone;
            for (int k = 0; k < 3; ++k) mt[k] = mt[k] + e->translation[k] * rate;
            const float ident[4] = {0, 0, 0, 1};
            float sl[4];
            q_slerp(ident, e->rotation, rate, sl);
            q_mul(mr, sl, mr);
        }
        break;
    default: /* UV, extra UV, material: no-op in libmmd (poser_impl.inl:355-358) */
        break;
    }
}

/* Poser::PrePhysicsPosing / PostPhysicsPosing, L/motion/poser_impl.inl:362-394 */
static void skin_matrices(state_t* s, const int32_t* list, uint32_t n) {
    const model_t* m = s->m;
    for (uint32_t i = 0; i < n; ++i) {
        int32_t b = list[i];
        float G[16];
        m_identity(G);
        for (int k = 0; k < 3; ++k) G[12 + k] = -m->bpos[3 * b + k];
        m_mul(G, s->local + 16 * b, s->skin + 16 * b);
    }
}
static void pre_physics(state_t* s) {
    const model_t* m = s->m;
    memset(s->vimg, 0, 12u * (size_t)m->nv);
    for (uint32_t b = 0; b < m->nb; ++b) {
        const float ident[4] = {0, 0, 0, 1};
        memset(s->morphT + 3 * b, 0, 12); memcpy(s->morphR + 4 * b, ident, 16);
        m_identity(s->local + 16 * b);
        memcpy(s->preIK + 4 * b, ident, 16); memcpy(s->ikR + 4 * b, ident, 16);
        memcpy(s->totR + 4 * b, ident, 16); memset(s->totT + 3 * b, 0, 12);
    }
    for (uint32_t i = 0; i < m->nm; ++i) update_morph(s, i, s->rate[i]);
    for (uint32_t i = 0; i < m->n_pre; ++i) update_bone(s, (uint32_t)m->order_pre[i]);
    skin_matrices(s, m->order_pre, m->n_pre);
}
static void post_physics(state_t* s) {
    const model_t* m = s->m;
    for (uint32_t i = 0; i < m->n_post; ++i) update_bone(s, (uint32_t)m->order_post[i]);
    skin_matrices(s, m->order_post, m->n_post);
}

/* Poser::Deform, L/motion/poser_impl.inl:396-461; transform / rotate math_impl.inl:1032-1045 */
static void deform(state_t* s) {
    const model_t* m = s->m;
    for (uint32_t i = 0; i < m->nv; ++i) {
        const int32_t* id = m->sid + 4 * i;
        const float* w = m->sw + 4 * i;
        float p[3], M[16];
        for (int k = 0; k < 3; ++k) p[k] = m->pos[3 * (size_t)i + k] + s->vimg[3 * (size_t)i + k];
        const float* n = m->nrm + 3 * (size_t)i;
        switch (m->stype[i]) {
        case MMDGPU_SKIN_BDEF1:
            memcpy(M, s->skin + 16 * id[0], 64);
            break;
        case MMDGPU_SKIN_BDEF4: {
            const float *m0 = s->skin + 16 * id[0], *m1 = s->skin + 16 * id[1], *m2 = s->skin + 16 * id[2],
                        *m3 = s->skin + 16 * id[3];
            for (int k = 0; k < 16; ++k) M[k] = m0[k] * w[0] + m1[k] * w[1] + m2[k] * w[2] + m3[k] * w[3];
            break;
        }
        default: { /* BDEF2, SDEF and anything else: Lerp(mat_1, mat_0)[w], math_impl.inl:1246-1254 */
            const float *m0 = s->skin + 16 * id[0], *m1 = s->skin + 16 * id[1];
            float l = w[0];
            if (l < EPS_F) memcpy(M, m1, 64);
            else if (l > (float)(1.0 - EPS_D)) memcpy(M, m0, 64);
            else for (int k = 0; k < 16; ++k) M[k] = (1.0f - l) * m1[k] + l * m0[k];
            break;
        }
        }
        float* op = s->opos + 3 * (size_t)i;
        float* on = s->onrm + 3 * (size_t)i;
        for (int k = 0; k < 3; ++k) {
            op[k] = p[0] * M[k] + p[1] * M[4 + k] + p[2] * M[8 + k] + M[12 + k];
            on[k] = n[0] * M[k] + n[1] * M[4 + k] + n[2] * M[8 + k];
        }
    }
}

/* main.cpp:1788-1821 with physics off.  ResetPosing's embedded Pre+Post evaluation is dead work
 * (every field it writes is rewritten by the PrePhysicsPosing that follows) and is not repeated. */
static void one_frame(state_t* s, uint32_t frame) {
    reset_poses(s);
    seek_frame(s, frame);
    pre_physics(s);
    post_physics(s);
    deform(s);
}

/* ---- exported interface (same shape as oracle/ref_harness.cc) ---------------------------------- */
EXPORT struct port_session* port_create(const mmdgpu_model_desc* md, const mmdgpu_anim_desc* ad) {
    struct port_session* p = calloc(1, sizeof *p);
    build_model(&p->m, md);
    if (ad) build_motion(&p->m, ad);
    state_init(&p->s, &p->m);
    return p;
}
EXPORT void port_destroy(struct port_session* p) {
    if (!p) return;
    state_free(&p->s);
    model_t* m = &p->m;
    void* ptrs[] = {m->pos, m->nrm, m->uv, m->stype, m->sid, m->sw, m->bpos, m->parent, m->level, m->flags,
                    m->has_parent, m->has_append, m->app_rot, m->app_trans, m->has_ik, m->is_link, m->app_parent,
                    m->app_ratio, m->local_offset, m->ik_target, m->ik_iters, m->ik_lbegin, m->ik_lcount,
                    m->ik_angle, m->l_bone, m->l_limited, m->l_fix, m->l_order, m->l_min, m->l_max, m->order_pre,
                    m->order_post, m->mtype, m->mbegin, m->mcount, m->ve, m->be, m->ge, m->bt_bone, m->mt_morph,
                    m->bt_begin, m->bt_count, m->mt_begin, m->mt_count, m->bkeys, m->mkeys};
    for (size_t i = 0; i < sizeof ptrs / sizeof ptrs[0]; ++i) free(ptrs[i]);
    free(p);
}
EXPORT void port_get_skinning(struct port_session* p, uint8_t* type, int32_t* id4, float* w4) {
    const model_t* m = &p->m;
    for (uint32_t i = 0; i < m->nv; ++i) {
        uint8_t t = m->stype[i];
        type[i] = t;
        int n = t == MMDGPU_SKIN_BDEF1 ? 1 : (t == MMDGPU_SKIN_BDEF4 ? 4 : 2);
        for (int k = 0; k < 4; ++k) {
            id4[4 * i + k] = k < n ? m->sid[4 * i + k] : -1;
            w4[4 * i + k] = 0;
        }
        if (t == MMDGPU_SKIN_BDEF4) for (int k = 0; k < 4; ++k) w4[4 * i + k] = m->sw[4 * i + k];
        else if (t != MMDGPU_SKIN_BDEF1) w4[4 * i] = m->sw[4 * i];
    }
}
EXPORT uint32_t port_get_ik_class(struct port_session* p, uint8_t* fix, uint8_t* order, uint32_t cap) {
    const model_t* m = &p->m;
    uint32_t o = 0;
    for (uint32_t b = 0; b < m->nb; ++b)
        if (m->has_ik[b])
            for (uint32_t j = 0; j < m->ik_lcount[b]; ++j) {
                if (o < cap) { fix[o] = m->l_fix[m->ik_lbegin[b] + j]; order[o] = m->l_order[m->ik_lbegin[b] + j]; }
                ++o;
            }
    return o;
}
static void copy_out(state_t* s, float* pos, float* nrm, float* skin, float* local, float* poses, float* rates) {
    const model_t* m = s->m;
    if (pos) memcpy(pos, s->opos, 12u * (size_t)m->nv);
    if (nrm) memcpy(nrm, s->onrm, 12u * (size_t)m->nv);
    if (skin) memcpy(skin, s->skin, 64u * m->nb);
    if (local) memcpy(local, s->local, 64u * m->nb);
    if (poses)
        for (uint32_t b = 0; b < m->nb; ++b) {
            memcpy(poses + 7 * b, s->T + 3 * b, 12);
            memcpy(poses + 7 * b + 3, s->R + 4 * b, 16);
        }
    if (rates) memcpy(rates, s->rate, 4u * m->nm);
}
EXPORT int port_run_frame(struct port_session* p, uint32_t frame, float* pos, float* nrm, float* skin, float* local,
                          float* poses, float* rates) {
    one_frame(&p->s, frame);
    copy_out(&p->s, pos, nrm, skin, local, poses, rates);
    return 0;
}
/* Host physics hand-back between Pre and Post (mmd-bullet_impl.inl:34-56): skinning / local matrices overwritten in place. */
EXPORT int port_run_frame_override(struct port_session* p, uint32_t frame, uint32_t n, const int32_t* bones, const float* skin16,
                                   const float* local16_or_null, float* pos, float* nrm, float* skin, float* local) {
    state_t* s = &p->s;
    reset_poses(s);
    seek_frame(s, frame);
    pre_physics(s);
    for (uint32_t i = 0; i < n; ++i) {
        memcpy(s->skin + 16 * (size_t)bones[i], skin16 + 16 * (size_t)i, 64);
        if (local16_or_null) memcpy(s->local + 16 * (size_t)bones[i], local16_or_null + 16 * (size_t)i, 64);
    }
    post_physics(s);
    deform(s);
    copy_out(s, pos, nrm, skin, local, NULL, NULL);
    return 0;
}
EXPORT int port_run_time(struct port_session* p, double seconds, float* pos, float* nrm, float* skin, float* poses,
                         float* rates) {
    state_t* s = &p->s;
    reset_poses(s);
    seek_time(s, seconds);
    pre_physics(s);
    post_physics(s);
    deform(s);
    copy_out(s, pos, nrm, skin, NULL, poses, rates);
    return 0;
}
EXPORT int port_run_manual(struct port_session* p, uint32_t nbp, const int32_t* bone, const float* pose7, uint32_t nmp,
                           const int32_t* morph, const float* weight, float* pos, float* nrm, float* skin) {
    state_t* s = &p->s;
    reset_poses(s);
    for (uint32_t i = 0; i < nbp; ++i) {
        memcpy(s->T + 3 * bone[i], pose7 + 7 * i, 12);
        memcpy(s->R + 4 * bone[i], pose7 + 7 * i + 3, 16);
    }
    for (uint32_t i = 0; i < nmp; ++i) s->rate[morph[i]] = weight[i];
    pre_physics(s);
    post_physics(s);
    deform(s);
    copy_out(s, pos, nrm, skin, NULL, NULL, NULL);
    return 0;
}
/* main.cpp:838-859 */
EXPORT void port_repack_sokol32(struct port_session* p, float* out8) {
    const model_t* m = &p->m;
    const state_t* s = &p->s;
    const float mmd_to_meter = 0.1f;
    for (uint32_t i = 0; i < m->nv; ++i) {
        float* o = out8 + 8 * (size_t)i;
        for (int k = 0; k < 3; ++k) { o[k] = s->opos[3 * (size_t)i + k] * mmd_to_meter; o[3 + k] = s->onrm[3 * (size_t)i + k]; }
        o[6] = m->uv[2 * (size_t)i]; o[7] = m->uv[2 * (size_t)i + 1];
    }
}

typedef struct { const model_t* m; const uint32_t* frames; uint32_t lo, hi; double acc; } job_t;
static void* job_main(void* arg) {
    job_t* j = (job_t*)arg;
    state_t s;
    state_init(&s, j->m);
    double acc = 0;
    for (uint32_t i = j->lo; i < j->hi; ++i) {
        one_frame(&s, j->frames[i]);
        if (j->m->nv) acc += s.opos[3 * (size_t)(i % j->m->nv)] + s.onrm[3 * (size_t)((i * 7u) % j->m->nv) + 1];
    }
    j->acc = acc;
    state_free(&s);
    return NULL;
}
EXPORT double port_time_frames(struct port_session* p, const uint32_t* frames, uint32_t n, uint32_t n_threads,
                               double* checksum) {
    if (n_threads < 1) n_threads = 1;
    job_t* jobs = calloc(n_threads, sizeof *jobs);
    pthread_t* th = calloc(n_threads, sizeof *th);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint32_t t = 0; t < n_threads; ++t) {
        jobs[t].m = &p->m; jobs[t].frames = frames;
        jobs[t].lo = (uint32_t)((uint64_t)n * t / n_threads);
        jobs[t].hi = (uint32_t)((uint64_t)n * (t + 1) / n_threads);
        if (n_threads == 1) job_main(&jobs[t]); else pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    double c = 0;
    for (uint32_t t = 0; t < n_threads; ++t) { if (n_threads > 1) pthread_join(th[t], NULL); c += jobs[t].acc; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) *checksum = c;
    free(jobs); free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}


/* ---- function-level known-answer interface (tests/test_math_kat.py): one row of `in` per case --------------------
 * op 0 BEZIER  in: 4 control bytes (as floats), x            out: lambda              (Bezier::SetC + operator[])
 * op 1 NLERP   in: a[4], b[4], l                             out: v[4]                (NLerpProxy<Vector4f>::operator[])
 * op 2 SLERP   in: a[4], b[4], l                             out: q[4]                (SLerpProxy<Quaternionf>::operator[])
 * op 3 Q2E     in: q[4], order (0 YZX, 1 ZXY, 2 XYZ)         out: euler[3]            (QuaternionTo{YZX,ZXY,XYZ})
 * op 4 E2Q     in: euler[3], order                           out: q[4]                ({YZX,ZXY,XYZ}ToQuaternion)
 * op 5 AXIS    in: axis[3], angle                            out: q[4]                (AxisToQuaternion)
 * op 6 QMUL    in: a[4], b[4]                                out: q[4]                (Quaternion::operator*)
 * op 7 QROT    in: q[4]                                      out: rows 0..2 (9)       (Quaternion::ToRotateMatrix)
 * op 8 QINV    in: q[4]                                      out: q[4]                (Quaternion::Inverse)
 * op 9 MATMUL  in: a[16], b[16]                              out: m[16]               (Matrix4x4::operator*)
 * op 10 VNORM  in: v[3]                                      out: v[3]                (Vector3D::Normalize)           */
static const int kat_in[11] = {5, 9, 9, 5, 4, 4, 8, 4, 4, 32, 3};
static const int kat_out[11] = {1, 4, 4, 3, 4, 4, 4, 9, 4, 16, 3};
EXPORT int port_math_kat(int op, const float* in, uint32_t n, float* out) {
    if (op < 0 || op > 10) return -1;
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + (size_t)i * kat_in[op];
        float* o = out + (size_t)i * kat_out[op];
        switch (op) {
        case 0: {
            bezier b;
            int8_t c[4] = {(int8_t)a[0], (int8_t)a[1], (int8_t)a[2], (int8_t)a[3]};
            bezier_set(&b, c);
            o[0] = bezier_at(&b, a[4]);
            break;
        }
        case 1: {
            const float* x = a; const float* y = a + 4; float l = a[8];
            if (l < EPS_F) { memcpy(o, x, 16); break; }
            if (l > (1.0f - EPS_F)) { memcpy(o, y, 16); break; }
            float dot = x[0] * y[0] + x[1] * y[1] + x[2] * y[2] + x[3] * y[3];
            float v[4];
            if (dot < 0.0f) for (int c = 0; c < 4; ++c) v[c] = (1.0f - l) * x[c] - l * y[c];
            else for (int c = 0; c < 4; ++c) v[c] = (1.0f - l) * x[c] + l * y[c];
            float nn = 1.0f / m_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
            for (int c = 0; c < 4; ++c) o[c] = v[c] * nn;
            break;
        }
        case 2: q_slerp(a, a + 4, a[8], o); break;
        case 3: quat_to_euler((int)a[4], a, o); break;
        case 4: euler_to_quat((int)a[3], a, o); break;
        case 5: axis_to_quat(a, a[3], o); break;
        case 6: q_mul(a, a + 4, o); break;
        case 7: { float m[16]; q_to_matrix(a, m); for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = m[4 * r + c]; break; }
        case 8: q_inverse(a, o); break;
        case 9: m_mul(a, a + 16, o); break;
        case 10: v3_normalize(a, o); break;
        }
    }
    return 0;
}
