// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked/imported by the product path (parry_b200/).
// CPU restatement of parry3d (f32, dim3) arithmetic. The linear algebra lives in nalgebra 0.34 /
// simba 0.9 which are NOT vendored under /root/reference (crates/parry3d/Cargo.toml:78-79, no
// Cargo.lock) => operation order below is the published nalgebra algorithm as recollected in
// SURVEY.md Appendix B ("parity unpinned" for the nalgebra layer).
// Compile with -ffp-contract=off: Rust never fuses a*b+c.
#pragma once
#include <cmath>
#include <cstdint>
#include <cfloat>
#include <algorithm>

namespace pb2o {

typedef float Real;
static const Real REAL_MAX = FLT_MAX;
static const Real DEFAULT_EPSILON = FLT_EPSILON;  // src/lib.rs:102

// Rust f32::min/max ignore a NaN operand; fminf/fmaxf have the same contract.
static inline Real rmin(Real a, Real b) { return fminf(a, b); }
static inline Real rmax(Real a, Real b) { return fmaxf(a, b); }

struct Vec3 {
    Real x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(Real x_, Real y_, Real z_) : x(x_), y(y_), z(z_) {}
    Real& operator[](int i) { return (&x)[i]; }
    const Real& operator[](int i) const { return (&x)[i]; }
};

static inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
static inline Vec3 operator*(const Vec3& a, Real s) { return Vec3(a.x * s, a.y * s, a.z * s); }
static inline Vec3 operator/(const Vec3& a, Real s) { return Vec3(a.x / s, a.y / s, a.z / s); }
static inline bool operator==(const Vec3& a, const Vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline bool operator!=(const Vec3& a, const Vec3& b) { return !(a == b); }

// nalgebra blas.rs `dotx` fixed-size U3 special case: a + b + c, left to right.
static inline Real dot(const Vec3& a, const Vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline Real norm_squared(const Vec3& a) { return dot(a, a); }
static inline Real norm(const Vec3& a) { return sqrtf(norm_squared(a)); }
// `normalize` = unscale by norm: component-wise DIVISION.
static inline Vec3 normalize(const Vec3& a) { return a / norm(a); }
static inline Vec3 vinf(const Vec3& a, const Vec3& b) { return Vec3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z)); }
static inline Vec3 vsup(const Vec3& a, const Vec3& b) { return Vec3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z)); }
// na::center(a, b) = (a + b) * 0.5  (nalgebra geometry/point_ops)
static inline Vec3 center(const Vec3& a, const Vec3& b) { return (a + b) * 0.5f; }
// imax: first index of the maximum
static inline int imax(const Vec3& v) {
    int best = 0;
    if (v.y > v[best]) best = 1;
    if (v.z > v[best]) best = 2;
    return best;
}

// Unit::try_new_and_get(v, min_norm)
static inline bool try_normalize_and_get(const Vec3& v, Real min_norm, Vec3& out, Real& n) {
    Real sq = norm_squared(v);
    if (sq > min_norm * min_norm) {
        n = sqrtf(sq);
        out = v / n;
        return true;
    }
    return false;
}
static inline bool try_normalize(const Vec3& v, Real min_norm, Vec3& out) {
    Real n;
    return try_normalize_and_get(v, min_norm, out, n);
}

// Unit quaternion stored (i, j, k, w) like nalgebra's coords.
struct Quat {
    Real i, j, k, w;
    Quat() : i(0), j(0), k(0), w(1) {}
    Quat(Real i_, Real j_, Real k_, Real w_) : i(i_), j(j_), k(k_), w(w_) {}
    Vec3 vec() const { return Vec3(i, j, k); }
    Quat conj() const { return Quat(-i, -j, -k, w); }
};

// UnitQuaternion * Vector3: t = (q.xyz × v) * 2; t * w + (q.xyz × t) + v
static inline Vec3 rotate(const Quat& q, const Vec3& v) {
    Vec3 t = cross(q.vec(), v) * 2.0f;
    Vec3 c = cross(q.vec(), t);
    return (t * q.w + c) + v;
}
static inline Vec3 inv_rotate(const Quat& q, const Vec3& v) { return rotate(q.conj(), v); }

// Hamilton product, nalgebra quaternion_ops.rs term order.
static inline Quat qmul(const Quat& a, const Quat& b) {
    return Quat(a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j,
                a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i,
                a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w,
                a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k);
}

struct Iso {
    Quat rot;
    Vec3 tra;
    Iso() {}
    Iso(const Quat& r, const Vec3& t) : rot(r), tra(t) {}
    static Iso from7(const float* p) { return Iso(Quat(p[0], p[1], p[2], p[3]), Vec3(p[4], p[5], p[6])); }
    Vec3 transform_point(const Vec3& p) const { return rotate(rot, p) + tra; }
    Vec3 transform_vector(const Vec3& v) const { return rotate(rot, v); }
    Vec3 inverse_transform_point(const Vec3& p) const { return inv_rotate(rot, p - tra); }
    Vec3 inverse_transform_vector(const Vec3& v) const { return inv_rotate(rot, v); }
    // a.inv_mul(b) (nalgebra isometry.rs): translation = Ra^-1 (tb - ta); rotation = Ra^-1 * Rb
    Iso inv_mul(const Iso& b) const {
        Quat inv = rot.conj();
        Vec3 tr12 = b.tra - tra;
        return Iso(qmul(inv, b.rot), rotate(inv, tr12));
    }
    // Isometry::inverse(): rotation^-1, translation = -(R^-1 t)  (nalgebra: inverse_mut)
    Iso inverse() const {
        Quat inv = rot.conj();
        Vec3 t = rotate(inv, tra);
        return Iso(inv, -t);
    }
    // utils/isometry_ops.rs:16-18: |R| * v with R = to_rotation_matrix()
    Vec3 absolute_transform_vector(const Vec3& v) const {
        Real i = rot.i, j = rot.j, k = rot.k, w = rot.w;
        Real ww = w * w, ii = i * i, jj = j * j, kk = k * k;
        Real ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f;
        Real ik = i * k * 2.0f, jk = j * k * 2.0f, wi = w * i * 2.0f;
        Real m00 = fabsf(ww + ii - jj - kk), m01 = fabsf(ij - wk), m02 = fabsf(wj + ik);
        Real m10 = fabsf(wk + ij), m11 = fabsf(ww - ii + jj - kk), m12 = fabsf(jk - wi);
        Real m20 = fabsf(ik - wj), m21 = fabsf(wi + jk), m22 = fabsf(ww - ii - jj + kk);
        // gemv by columns: res = col0*v0; res += col1*v1; res += col2*v2
        return Vec3((m00 * v.x + m01 * v.y) + m02 * v.z,
                    (m10 * v.x + m11 * v.y) + m12 * v.z,
                    (m20 * v.x + m21 * v.y) + m22 * v.z);
    }
};

// approx::relative_eq! defaults (epsilon = max_relative = f32::EPSILON)
static inline bool relative_eq(Real a, Real b, Real eps = FLT_EPSILON, Real max_rel = FLT_EPSILON) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    Real d = fabsf(a - b);
    if (d <= eps) return true;
    Real aa = fabsf(a), ab = fabsf(b);
    Real largest = ab > aa ? ab : aa;
    return d <= largest * max_rel;
}

struct Aabb {
    Vec3 mins, maxs;
    Aabb() {}
    Aabb(const Vec3& a, const Vec3& b) : mins(a), maxs(b) {}
    static Aabb new_invalid() { return Aabb(Vec3(REAL_MAX, REAL_MAX, REAL_MAX), Vec3(-REAL_MAX, -REAL_MAX, -REAL_MAX)); }
    Vec3 center() const { return pb2o::center(mins, maxs); }
    Vec3 extents() const { return maxs - mins; }
    Real volume() const { Vec3 e = extents(); return e.x * e.y * e.z; }                 // aabb.rs:448-454
    Real half_area() const { Vec3 e = extents(); return e.x * (e.y + e.z) + e.y * e.z; } // aabb.rs:473-476
    void merge(const Aabb& o) { mins = vinf(mins, o.mins); maxs = vsup(maxs, o.maxs); }
    Aabb merged(const Aabb& o) const { return Aabb(vinf(mins, o.mins), vsup(maxs, o.maxs)); }
    // aabb.rs:951-953 inclusive on all axes (na::partial_le / partial_ge: all components)
    bool intersects(const Aabb& o) const {
        return mins.x <= o.maxs.x && mins.y <= o.maxs.y && mins.z <= o.maxs.z &&
               maxs.x >= o.mins.x && maxs.y >= o.mins.y && maxs.z >= o.mins.z;
    }
    bool contains(const Aabb& o) const {
        return mins.x <= o.mins.x && mins.y <= o.mins.y && mins.z <= o.mins.z &&
               maxs.x >= o.maxs.x && maxs.y >= o.maxs.y && maxs.z >= o.maxs.z;
    }
    bool contains_local_point(const Vec3& p) const {
        return mins.x <= p.x && mins.y <= p.y && mins.z <= p.z && p.x <= maxs.x && p.y <= maxs.y && p.z <= maxs.z;
    }
};

struct Ray {
    Vec3 origin, dir;
    Ray() {}
    Ray(const Vec3& o, const Vec3& d) : origin(o), dir(d) {}
    // ray.rs:167-172
    Ray inverse_transform_by(const Iso& m) const {
        return Ray(m.inverse_transform_point(origin), m.inverse_transform_vector(dir));
    }
    Vec3 point_at(Real t) const { return origin + dir * t; }
};

}  // namespace pb2o
