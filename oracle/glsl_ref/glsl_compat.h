// glsl_compat.h — GLSL 3.30 vocabulary for g++.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Lets the reference's OWN shader text (src/shaders/*.glsl, read where it lies under /root/reference and
// passed through oracle/glsl_ref/glsl2cpp.py) compile and execute as C++: vector/matrix types with
// swizzles, the built-ins the shaders call, and samplers over plain host arrays.
//
// Included INSIDE a namespace (once per shader program) after the system headers, so that the
// unqualified names the shader text uses (sqrt, pow, min, max, mix ...) resolve to the fp32 definitions
// below and never to libm's double overloads.
//
// Arithmetic: every operator is one IEEE fp32 operation per component (the TU is built with
// -ffp-contract=off).  Built-ins GLSL leaves implementation-defined are pinned to their spec formula:
//   normalize(v) = v / sqrt(dot(v,v));   mix(a,b,t) = a*(1-t) + b*t;   reflect / refract = GLSL 4.60 §8.5;
//   inverse() = adjugate / determinant;   matrix*vector sums the columns left to right;
//   texture() on a LINEAR sampler = GL 4.6 §8.14.2 bilinear with fp32 weights, REPEAT wrap;
//   texelFetch on RGB32F = (r,g,b,1);   float->int conversions truncate.

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct ivec4;

// ---- swizzle views: a window over the N components of the owning vector ---------------------------------------
template <class VT, class T, int N, int A, int B> struct swz2
{
    T d[N];
    operator VT() const { return VT(d[A], d[B]); }
    swz2& operator=(const VT& v) { T a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
    swz2& operator=(const swz2& o) { return *this = VT(o); }
    swz2& operator+=(const VT& v) { return *this = VT(*this) + v; }
    swz2& operator-=(const VT& v) { return *this = VT(*this) - v; }
    swz2& operator*=(const VT& v) { return *this = VT(*this) * v; }
    swz2& operator/=(const VT& v) { return *this = VT(*this) / v; }
};
template <class VT, class T, int N, int A, int B, int C> struct swz3
{
    T d[N];
    operator VT() const { return VT(d[A], d[B], d[C]); }
    swz3& operator=(const VT& v) { T a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }
    swz3& operator=(const swz3& o) { return *this = VT(o); }
    swz3& operator+=(const VT& v) { return *this = VT(*this) + v; }
    swz3& operator-=(const VT& v) { return *this = VT(*this) - v; }
    swz3& operator*=(const VT& v) { return *this = VT(*this) * v; }
    swz3& operator/=(const VT& v) { return *this = VT(*this) / v; }
};
template <class VT, class T, int N, int A, int B, int C, int D> struct swz4
{
    T d[N];
    operator VT() const { return VT(d[A], d[B], d[C], d[D]); }
    swz4& operator=(const VT& v) { T a = v.x, b = v.y, c = v.z, e = v.w; d[A] = a; d[B] = b; d[C] = c; d[D] = e; return *this; }
    swz4& operator=(const swz4& o) { return *this = VT(o); }
};

// ---- float vectors ---------------------------------------------------------------------------------------------
struct vec2
{
    union {
        struct { float x, y; };
        struct { float r, g; };
        swz2<vec2, float, 2, 0, 1> xy;
        swz2<vec2, float, 2, 1, 0> yx;
    };
    vec2() { x = 0.0f; y = 0.0f; }
    explicit vec2(float s) { x = s; y = s; }
    vec2(float a, float b) { x = a; y = b; }
    vec2(int a, int b) { x = (float)a; y = (float)b; }
    explicit vec2(const ivec2& v);
    explicit vec2(const vec3& v);
    vec2(const vec2& o) { x = o.x; y = o.y; }
    vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
};
struct vec3
{
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz2<vec2, float, 3, 0, 1> xy;
        swz2<vec2, float, 3, 1, 2> yz;
        swz2<vec2, float, 3, 0, 2> xz;
        swz3<vec3, float, 3, 0, 1, 2> xyz;
        swz3<vec3, float, 3, 0, 1, 2> rgb;
    };
    vec3() { x = 0.0f; y = 0.0f; z = 0.0f; }
    explicit vec3(float s) { x = s; y = s; z = s; }
    vec3(float a, float b, float c) { x = a; y = b; z = c; }
    vec3(const vec2& v, float c) { x = v.x; y = v.y; z = c; }
    vec3(const vec2& v, int c) { x = v.x; y = v.y; z = (float)c; }
    vec3(float a, const vec2& v) { x = a; y = v.x; z = v.y; }
    explicit vec3(const vec4& v);
    explicit vec3(const ivec3& v);
    vec3(const vec3& o) { x = o.x; y = o.y; z = o.z; }
    vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
struct vec4
{
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<vec2, float, 4, 0, 1> xy;
        swz2<vec2, float, 4, 2, 3> zw;
        swz2<vec2, float, 4, 0, 1> rg;
        swz2<vec2, float, 4, 2, 1> bg;
        swz3<vec3, float, 4, 0, 1, 2> xyz;
        swz3<vec3, float, 4, 0, 1, 2> rgb;
        swz3<vec3, float, 4, 3, 0, 1> wxy;
        swz4<vec4, float, 4, 0, 1, 2, 3> xyzw;
        swz4<vec4, float, 4, 0, 1, 2, 3> rgba;
    };
    vec4() { x = 0.0f; y = 0.0f; z = 0.0f; w = 0.0f; }
    explicit vec4(float s) { x = s; y = s; z = s; w = s; }
    vec4(float a, float b, float c, float d) { x = a; y = b; z = c; w = d; }
    vec4(const vec3& v, float d) { x = v.x; y = v.y; z = v.z; w = d; }
    vec4(const vec2& v, float c, float d) { x = v.x; y = v.y; z = c; w = d; }
    vec4(const vec2& u, const vec2& v) { x = u.x; y = u.y; z = v.x; w = v.y; }
    explicit vec4(const ivec4& v);
    vec4(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; }
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
};
inline vec2::vec2(const vec3& v) { x = v.x; y = v.y; }
inline vec3::vec3(const vec4& v) { x = v.x; y = v.y; z = v.z; }

// ---- integer vectors -------------------------------------------------------------------------------------------
struct ivec2
{
    union { struct { int x, y; }; swz2<ivec2, int, 2, 0, 1> xy; };
    ivec2() { x = 0; y = 0; }
    explicit ivec2(int s) { x = s; y = s; }
    ivec2(int a, int b) { x = a; y = b; }
    explicit ivec2(const vec2& v) { x = (int)v.x; y = (int)v.y; }
    ivec2(const ivec2& o) { x = o.x; y = o.y; }
    ivec2& operator=(const ivec2& o) { x = o.x; y = o.y; return *this; }
};
struct ivec3
{
    union { struct { int x, y, z; }; swz3<ivec3, int, 3, 0, 1, 2> xyz; swz2<ivec2, int, 3, 0, 1> xy; };
    ivec3() { x = 0; y = 0; z = 0; }
    explicit ivec3(int s) { x = s; y = s; z = s; }
    ivec3(int a, int b, int c) { x = a; y = b; z = c; }
    explicit ivec3(const vec3& v) { x = (int)v.x; y = (int)v.y; z = (int)v.z; }
    ivec3(const ivec3& o) { x = o.x; y = o.y; z = o.z; }
    ivec3& operator=(const ivec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
struct ivec4
{
    union { struct { int x, y, z, w; }; swz3<ivec3, int, 4, 0, 1, 2> xyz; swz2<ivec2, int, 4, 0, 1> xy; };
    ivec4() { x = 0; y = 0; z = 0; w = 0; }
    explicit ivec4(int s) { x = s; y = s; z = s; w = s; }
    ivec4(int a, int b, int c, int d) { x = a; y = b; z = c; w = d; }
    explicit ivec4(const vec4& v) { x = (int)v.x; y = (int)v.y; z = (int)v.z; w = (int)v.w; }
    ivec4(const ivec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; }
    ivec4& operator=(const ivec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
};
inline vec2::vec2(const ivec2& v) { x = (float)v.x; y = (float)v.y; }
inline vec3::vec3(const ivec3& v) { x = (float)v.x; y = (float)v.y; z = (float)v.z; }
inline vec4::vec4(const ivec4& v) { x = (float)v.x; y = (float)v.y; z = (float)v.z; w = (float)v.w; }

struct uvec4
{
    uint x, y, z, w;
    uvec4() { x = 0u; y = 0u; z = 0u; w = 0u; }
    uvec4(uint a, uint b, uint c, uint d) { x = a; y = b; z = c; w = d; }
    uvec4(const vec2& p, uint c, uint d) { x = (uint)p.x; y = (uint)p.y; z = c; w = d; }
};
inline uvec4 operator*(uvec4 a, uint s) { return uvec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline uvec4 operator+(uvec4 a, uint s) { return uvec4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline uvec4 operator>>(uvec4 a, uint s) { return uvec4(a.x >> s, a.y >> s, a.z >> s, a.w >> s); }
inline uvec4 operator^(uvec4 a, uvec4 b) { return uvec4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }

struct bvec3 { bool x, y, z; };
struct bvec4 { bool x, y, z, w; };
inline bool all(bvec3 b) { return b.x && b.y && b.z; }
inline bool all(bvec4 b) { return b.x && b.y && b.z && b.w; }
inline bool any(bvec3 b) { return b.x || b.y || b.z; }
inline bool any(bvec4 b) { return b.x || b.y || b.z || b.w; }
inline bvec4 greaterThanEqual(vec4 a, vec4 b) { return {a.x >= b.x, a.y >= b.y, a.z >= b.z, a.w >= b.w}; }
inline bvec3 greaterThanEqual(vec3 a, vec3 b) { return {a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bvec4 lessThan(vec4 a, vec4 b) { return {a.x < b.x, a.y < b.y, a.z < b.z, a.w < b.w}; }
inline bvec3 lessThan(vec3 a, vec3 b) { return {a.x < b.x, a.y < b.y, a.z < b.z}; }

// ---- component-wise operators (vector op vector, vector op scalar, scalar op vector) ----------------------------
#define GLSL_VEC_OPS(V, ...)                                                                      \
    inline V operator+(V a, V b) { return GLSL_MAP2(V, +, __VA_ARGS__); }                        \
    inline V operator-(V a, V b) { return GLSL_MAP2(V, -, __VA_ARGS__); }                        \
    inline V operator*(V a, V b) { return GLSL_MAP2(V, *, __VA_ARGS__); }                        \
    inline V operator/(V a, V b) { return GLSL_MAP2(V, /, __VA_ARGS__); }
#define GLSL_MAP2(V, op, ...) GLSL_PICK(__VA_ARGS__, GLSL_M4, GLSL_M3, GLSL_M2)(V, op)
#define GLSL_PICK(_1, _2, _3, _4, NAME, ...) NAME
#define GLSL_M2(V, op) V(a.x op b.x, a.y op b.y)
#define GLSL_M3(V, op) V(a.x op b.x, a.y op b.y, a.z op b.z)
#define GLSL_M4(V, op) V(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
GLSL_VEC_OPS(vec2, x, y)
GLSL_VEC_OPS(vec3, x, y, z)
GLSL_VEC_OPS(vec4, x, y, z, w)
GLSL_VEC_OPS(ivec2, x, y)
GLSL_VEC_OPS(ivec3, x, y, z)
#define GLSL_SCALAR_OPS(V)                                                                        \
    inline V operator+(V a, float s) { return a + V(s); }                                         \
    inline V operator-(V a, float s) { return a - V(s); }                                         \
    inline V operator*(V a, float s) { return a * V(s); }                                         \
    inline V operator/(V a, float s) { return a / V(s); }                                         \
    inline V operator+(float s, V a) { return V(s) + a; }                                         \
    inline V operator-(float s, V a) { return V(s) - a; }                                         \
    inline V operator*(float s, V a) { return V(s) * a; }                                         \
    inline V operator/(float s, V a) { return V(s) / a; }                                         \
    inline V& operator+=(V& a, V b) { a = a + b; return a; }                                      \
    inline V& operator-=(V& a, V b) { a = a - b; return a; }                                      \
    inline V& operator*=(V& a, V b) { a = a * b; return a; }                                      \
    inline V& operator/=(V& a, V b) { a = a / b; return a; }                                      \
    inline V& operator+=(V& a, float s) { a = a + s; return a; }                                  \
    inline V& operator-=(V& a, float s) { a = a - s; return a; }                                  \
    inline V& operator*=(V& a, float s) { a = a * s; return a; }                                  \
    inline V& operator/=(V& a, float s) { a = a / s; return a; }
GLSL_SCALAR_OPS(vec2)
GLSL_SCALAR_OPS(vec3)
GLSL_SCALAR_OPS(vec4)
inline vec2 operator-(vec2 a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(vec4 a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(vec3 a, vec3 b) { return !(a == b); }

// ---- scalar built-ins ----------------------------------------------------------------------------------------------
inline float sqrt(float x) { return ::sqrtf(x); }
inline float inversesqrt(float x) { return 1.0f / ::sqrtf(x); }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float asin(float x) { return ::asinf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float atan(float y, float x) { return ::atan2f(y, x); }
inline float atan(float x) { return ::atanf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float exp2(float x) { return ::exp2f(x); }
inline float log2(float x) { return ::log2f(x); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float fract(float x) { return x - ::floorf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline int   abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline int   min(int a, int b) { return b < a ? b : a; }
inline int   max(int a, int b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int   clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) { float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
inline bool  isnan(float x) { return x != x; }
inline bool  isinf(float x) { return ::fabsf(x) > 3.402823466e38f; }

// ---- vector built-ins ------------------------------------------------------------------------------------------------
#define GLSL_LIFT1(fn)                                                                            \
    inline vec2 fn(vec2 a) { return vec2(fn(a.x), fn(a.y)); }                                     \
    inline vec3 fn(vec3 a) { return vec3(fn(a.x), fn(a.y), fn(a.z)); }                            \
    inline vec4 fn(vec4 a) { return vec4(fn(a.x), fn(a.y), fn(a.z), fn(a.w)); }
GLSL_LIFT1(sqrt) GLSL_LIFT1(sin) GLSL_LIFT1(cos) GLSL_LIFT1(exp) GLSL_LIFT1(log) GLSL_LIFT1(floor) GLSL_LIFT1(fract)
GLSL_LIFT1(abs) GLSL_LIFT1(sign)
#define GLSL_LIFT2(fn)                                                                            \
    inline vec2 fn(vec2 a, vec2 b) { return vec2(fn(a.x, b.x), fn(a.y, b.y)); }                   \
    inline vec3 fn(vec3 a, vec3 b) { return vec3(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z)); }     \
    inline vec4 fn(vec4 a, vec4 b) { return vec4(fn(a.x, b.x), fn(a.y, b.y), fn(a.z, b.z), fn(a.w, b.w)); } \
    inline vec2 fn(vec2 a, float b) { return fn(a, vec2(b)); }                                    \
    inline vec3 fn(vec3 a, float b) { return fn(a, vec3(b)); }                                    \
    inline vec4 fn(vec4 a, float b) { return fn(a, vec4(b)); }
GLSL_LIFT2(min) GLSL_LIFT2(max) GLSL_LIFT2(pow) GLSL_LIFT2(mod)
inline vec2 clamp(vec2 v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec3 clamp(vec3 v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec4 clamp(vec4 v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec3 clamp(vec3 v, vec3 lo, vec3 hi) { return min(max(v, lo), hi); }
inline vec2 mix(vec2 a, vec2 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec4 mix(vec4 a, vec4 b, float t) { return a * (1.0f - t) + b * t; }
inline vec2 mix(vec2 a, vec2 b, vec2 t) { return a * (vec2(1.0f) - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, vec3 t) { return a * (vec3(1.0f) - t) + b * t; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float length(vec2 a) { return sqrt(dot(a, a)); }
inline float length(vec3 a) { return sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline vec2 normalize(vec2 a) { return a / sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / sqrt(dot(a, a)); }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }
inline vec3 refract(vec3 I, vec3 N, float eta)
{
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return vec3(0.0f);
    return I * eta - N * (eta * d + sqrt(k));
}

// ---- matrices (column-major, m[c] = column c, as GLSL) -----------------------------------------------------------
struct mat4;
struct mat3
{
    vec3 c[3];
    mat3() {}
    mat3(vec3 a, vec3 b, vec3 d) { c[0] = a; c[1] = b; c[2] = d; }
    explicit mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4
{
    vec4 c[4];
    mat4() {}
    mat4(vec4 a, vec4 b, vec4 d, vec4 e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m) { c[0] = vec3(m.c[0]); c[1] = vec3(m.c[1]); c[2] = vec3(m.c[2]); }
// M * v = v.x*col0 + v.y*col1 + ... ;  v * M = (dot(v,col0), dot(v,col1), ...)
inline vec3 operator*(const mat3& m, vec3 v)
{
    return vec3(v.x * m.c[0].x + v.y * m.c[1].x + v.z * m.c[2].x, v.x * m.c[0].y + v.y * m.c[1].y + v.z * m.c[2].y,
                v.x * m.c[0].z + v.y * m.c[1].z + v.z * m.c[2].z);
}
inline vec3 operator*(vec3 v, const mat3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
inline vec4 operator*(const mat4& m, vec4 v)
{
    return vec4(v.x * m.c[0].x + v.y * m.c[1].x + v.z * m.c[2].x + v.w * m.c[3].x, v.x * m.c[0].y + v.y * m.c[1].y + v.z * m.c[2].y + v.w * m.c[3].y,
                v.x * m.c[0].z + v.y * m.c[1].z + v.z * m.c[2].z + v.w * m.c[3].z, v.x * m.c[0].w + v.y * m.c[1].w + v.z * m.c[2].w + v.w * m.c[3].w);
}
inline mat3 transpose(const mat3& m)
{
    return mat3(vec3(m.c[0].x, m.c[1].x, m.c[2].x), vec3(m.c[0].y, m.c[1].y, m.c[2].y), vec3(m.c[0].z, m.c[1].z, m.c[2].z));
}
// inverse = adjugate / determinant.  Element (r, k) below means column k, component r.
inline mat3 inverse(const mat3& m)
{
    // treat the matrix as A[i][j] = m.c[i][j]; the inverse of the transpose is the transpose of the inverse, so the
    // same cofactor layout serves either convention.
    float a00 = m.c[0].x, a01 = m.c[0].y, a02 = m.c[0].z, a10 = m.c[1].x, a11 = m.c[1].y, a12 = m.c[1].z, a20 = m.c[2].x, a21 = m.c[2].y, a22 = m.c[2].z;
    float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    float det = a00 * c00 + a01 * c01 + a02 * c02;
    float id = 1.0f / det;
    return mat3(vec3(c00 * id, (a02 * a21 - a01 * a22) * id, (a01 * a12 - a02 * a11) * id),
                vec3(c01 * id, (a00 * a22 - a02 * a20) * id, (a02 * a10 - a00 * a12) * id),
                vec3(c02 * id, (a01 * a20 - a00 * a21) * id, (a00 * a11 - a01 * a10) * id));
}
inline mat4 inverse(const mat4& m)
{
    float a00 = m.c[0].x, a01 = m.c[0].y, a02 = m.c[0].z, a03 = m.c[0].w, a10 = m.c[1].x, a11 = m.c[1].y, a12 = m.c[1].z, a13 = m.c[1].w;
    float a20 = m.c[2].x, a21 = m.c[2].y, a22 = m.c[2].z, a23 = m.c[2].w, a30 = m.c[3].x, a31 = m.c[3].y, a32 = m.c[3].z, a33 = m.c[3].w;
    float s0 = a00 * a11 - a10 * a01, s1 = a00 * a12 - a10 * a02, s2 = a00 * a13 - a10 * a03;
    float s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    float c5 = a22 * a33 - a32 * a23, c4 = a21 * a33 - a31 * a23, c3 = a21 * a32 - a31 * a22;
    float c2 = a20 * a33 - a30 * a23, c1 = a20 * a32 - a30 * a22, c0 = a20 * a31 - a30 * a21;
    float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    float id = 1.0f / det;
    mat4 r;
    r.c[0] = vec4((a11 * c5 - a12 * c4 + a13 * c3) * id, (-a01 * c5 + a02 * c4 - a03 * c3) * id, (a31 * s5 - a32 * s4 + a33 * s3) * id, (-a21 * s5 + a22 * s4 - a23 * s3) * id);
    r.c[1] = vec4((-a10 * c5 + a12 * c2 - a13 * c1) * id, (a00 * c5 - a02 * c2 + a03 * c1) * id, (-a30 * s5 + a32 * s2 - a33 * s1) * id, (a20 * s5 - a22 * s2 + a23 * s1) * id);
    r.c[2] = vec4((a10 * c4 - a11 * c2 + a13 * c0) * id, (-a00 * c4 + a01 * c2 - a03 * c0) * id, (a30 * s4 - a31 * s2 + a33 * s0) * id, (-a20 * s4 + a21 * s2 - a23 * s0) * id);
    r.c[3] = vec4((-a10 * c3 + a11 * c1 - a12 * c0) * id, (a00 * c3 - a01 * c1 + a02 * c0) * id, (-a30 * s3 + a31 * s1 - a32 * s0) * id, (a20 * s3 - a21 * s1 + a22 * s0) * id);
    return r;
}

// ---- samplers over host arrays -----------------------------------------------------------------------------------
struct samplerBuffer  { const float* data = nullptr; int channels = 3; };        // GL_RGB32F / GL_RGBA32F texture buffers
struct isamplerBuffer { const int* data = nullptr; int channels = 3; };          // GL_RGB32I
struct sampler2D      { const float* data = nullptr; int w = 0, h = 0, channels = 4; bool linear = false; };
struct sampler2DArray { const unsigned char* data = nullptr; int w = 0, h = 0, layers = 0; };   // GL_RGBA8, LINEAR, REPEAT

inline vec4 glsl_texel(const float* p, int ch)
{
    return vec4(p[0], ch > 1 ? p[1] : 0.0f, ch > 2 ? p[2] : 0.0f, ch > 3 ? p[3] : 1.0f);
}
inline vec4 texelFetch(const samplerBuffer& s, int i) { return glsl_texel(s.data + (size_t)i * s.channels, s.channels); }
inline ivec4 texelFetch(const isamplerBuffer& s, int i)
{
    const int* p = s.data + (size_t)i * s.channels;
    return ivec4(p[0], p[1], p[2], s.channels > 3 ? p[3] : 1);
}
inline vec4 texelFetch(const sampler2D& s, ivec2 p, int /*lod*/) { return glsl_texel(s.data + ((size_t)p.y * s.w + p.x) * s.channels, s.channels); }
inline int glsl_wrap_repeat(float f, int n) { int i = (int)::fmodf(f, (float)n); if (i < 0) i += n; return i; }
inline vec4 texture(const sampler2D& s, vec2 uv)
{
    if (!s.linear)
    {   // NEAREST: texel containing the coordinate
        int x = glsl_wrap_repeat(::floorf(uv.x * (float)s.w), s.w), y = glsl_wrap_repeat(::floorf(uv.y * (float)s.h), s.h);
        return glsl_texel(s.data + ((size_t)y * s.w + x) * s.channels, s.channels);
    }
    float x = uv.x * (float)s.w - 0.5f, y = uv.y * (float)s.h - 0.5f;
    float fx0 = ::floorf(x), fy0 = ::floorf(y);
    float ax = x - fx0, ay = y - fy0;
    int x0 = glsl_wrap_repeat(fx0, s.w), x1 = glsl_wrap_repeat(fx0 + 1.0f, s.w), y0 = glsl_wrap_repeat(fy0, s.h), y1 = glsl_wrap_repeat(fy0 + 1.0f, s.h);
    vec4 t00 = glsl_texel(s.data + ((size_t)y0 * s.w + x0) * s.channels, s.channels), t10 = glsl_texel(s.data + ((size_t)y0 * s.w + x1) * s.channels, s.channels);
    vec4 t01 = glsl_texel(s.data + ((size_t)y1 * s.w + x0) * s.channels, s.channels), t11 = glsl_texel(s.data + ((size_t)y1 * s.w + x1) * s.channels, s.channels);
    return mix(mix(t00, t10, ax), mix(t01, t11, ax), ay);
}
inline vec4 texture(const sampler2DArray& s, vec3 uvl)
{
    int layer = (int)::floorf(uvl.z + 0.5f); layer = max(0, min(layer, s.layers - 1));
    float x = uvl.x * (float)s.w - 0.5f, y = uvl.y * (float)s.h - 0.5f;
    float fx0 = ::floorf(x), fy0 = ::floorf(y);
    float ax = x - fx0, ay = y - fy0;
    int x0 = glsl_wrap_repeat(fx0, s.w), x1 = glsl_wrap_repeat(fx0 + 1.0f, s.w), y0 = glsl_wrap_repeat(fy0, s.h), y1 = glsl_wrap_repeat(fy0 + 1.0f, s.h);
    const unsigned char* base = s.data + (size_t)layer * s.w * s.h * 4;
    auto tx = [&](int xx, int yy) {
        const unsigned char* p = base + ((size_t)yy * s.w + xx) * 4;
        return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f);
    };
    return mix(mix(tx(x0, y0), tx(x1, y0), ax), mix(tx(x0, y1), tx(x1, y1), ax), ay);
}
