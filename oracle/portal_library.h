// CPU restatement of the reference's predefined GLSL library.
//
// ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// parity unpinned: the reference holds no tests, golden vectors or fixtures for this path (oracle/README.md).
// Every function cites the lines of /root/reference/src/library.glsl it follows.
// Included by the per-scene translation unit that oracle/gen_oracle.py writes, after
// glsl_compat.h and after the uniform block (`PE_R` = renderer uniforms) is declared.
#pragma once

namespace pe_oracle {

// library.glsl:15-16
#define PI acos(PE_L(-1.))
#define PI2 (acos(PE_L(-1.)) / PE_L(2.0))

// library.glsl:19-21
static inline bool between(real a, real x, real b) { return a <= x && x <= b; }
// library.glsl:24-26
static inline real sqr(real a) { return a * a; }
// library.glsl:32-34
static inline vec3 sqrvec(vec3 v) { return vec3(sqr(v.x), sqr(v.y), sqr(v.z)); }

// library.glsl:40-46
struct Ray {
    vec4 o;
    vec4 d;
    real tmul;
    bool in_subspace;
};

// library.glsl:48-51
static inline Ray offset_ray(Ray r, real t) {
    r.o += r.d * t;
    return r;
}

// library.glsl:53
static const Ray ray_none = Ray{vec4(real(0)), vec4(real(0)), real(0), false};

// library.glsl:56-62
static inline vec3 normalize_normal(vec3 normal, vec3 dir) {
    normal = normalize(normal);
    if (dot(normal, dir) > real(0)) {
        normal *= real(-1);
    }
    return normal;
}

// library.glsl:65-67
static inline bool is_collinear(vec3 a, vec3 b) {
    return abs(dot(a, b) / (length(a) * length(b)) - real(1)) < PE_L(0.01);
}

// library.glsl:70-72
static inline vec3 my_reflect(vec3 dir, vec3 normal) {
    return dir - normal * dot(dir, normal) / dot(normal, normal) * real(2);
}

// library.glsl:75-92
static inline vec3 my_refract(vec3 dir, vec3 normal, real refractive_index) {
    real ri = refractive_index;
    bool from_outside = dot(normal, dir) > real(0);
    if (!from_outside) {
        ri = real(1) / ri;
    } else {
        normal = -normal;
    }
    dir = normalize(dir);
    real c = -dot(normal, dir);
    real d = PE_L(1.0) - ri * ri * (PE_L(1.0) - c * c);
    if (d > real(0)) {
        return dir * ri + normal * (ri * c - sqrt(d));
    } else {
        return my_reflect(dir, normal);
    }
}

// library.glsl:95-102
static inline Ray transform(const mat4& matrix, Ray r) {
    return Ray{matrix * r.o, matrix * r.d, r.tmul, r.in_subspace};
}

// library.glsl:104-106
static inline vec3 get_normal(const mat4& matrix) {
    vec4 v = matrix * vec4(real(0), real(0), real(1), real(0));
    return vec3(v.x, v.y, v.z);
}

// library.glsl:108-113
static inline Ray normalize_ray(Ray r) {
    real len = length(r.d);
    r.d /= len;
    r.tmul /= len;
    return r;
}

// library.glsl:116-120
static inline mat3 adjugate(const mat4& m) {
    return mat3(cross(vec3(m[1]), vec3(m[2])), cross(vec3(m[2]), vec3(m[0])), cross(vec3(m[0]), vec3(m[1])));
}

// library.glsl:127-133
struct SurfaceIntersection {
    bool hit;
    real t;
    real u;
    real v;
    vec3 n;
};

// library.glsl:136
static const SurfaceIntersection intersection_none = SurfaceIntersection{false, PE_L(1e10), real(0), real(0), vec3(real(0))};

// library.glsl:138-146
static inline SurfaceIntersection plane_intersect_normalized(Ray r) {
    real t = -r.o.z / r.d.z;
    if (t < real(0)) {
        return intersection_none;
    } else {
        vec4 pos = r.o + r.d * t;
        return SurfaceIntersection{true, t, pos.x, pos.y, vec3(real(0), real(0), real(1))};
    }
}

// library.glsl:149-162
static inline SurfaceIntersection plane_intersect(Ray r, const mat4& plane_inv, vec3 normal) {
    normal = normalize_normal(normal, vec3(r.d));
    r = transform(plane_inv, r);
    real len = length(r.d);
    r.d = normalize(r.d);

    SurfaceIntersection result = plane_intersect_normalized(r);
    if (result.hit) {
        result.t /= len;
        result.n = normal;
    }
    return result;
}

// library.glsl:169-171
static inline vec3 color(real r, real g, real b) { return vec3(r * r, g * g, b * b); }

// library.glsl:177-181
static inline real color_normal(vec3 normal, vec4 direction) {
    if (_angle_color_disable == 1) return PE_L(1.0);
    return abs(dot(normalize(vec3(direction)), normalize(normal)));
}

// library.glsl:184-188
static inline vec3 color_grid(vec3 start, vec2 uv) {
    if (_grid_disable == 1) return start;
    uv = fract(uv * PE_L(0.25));
    return start * mix(mix(PE_L(0.7), PE_L(1.1), step(uv.x, PE_L(0.5))), mix(PE_L(1.1), PE_L(0.7), step(uv.x, PE_L(0.5))),
                       step(uv.y, PE_L(0.5)));
}

// library.glsl:199-205
static inline real circle_sdf(vec2 position) {
    vec2 s = vec2(PE_L(2.0), sqrt(PE_L(3.0)) * PE_L(2.0));
    position /= s;
    vec2 d1 = (fract(position) - PE_L(0.5)) * s;
    vec2 d2 = (fract(position + PE_L(0.5)) - PE_L(0.5)) * s;
    return sqrt(min(dot(d1, d1), dot(d2, d2))) - PE_L(1.0);
}
// library.glsl:206-211
static inline vec3 color_grid2(vec3 start, vec2 uv) {
    real d = circle_sdf(uv);
    real val = PE_L(0.7);
    if (d < PE_L(-0.2)) val = PE_L(1.1);
    return start * val;
}

// library.glsl:268-283
static inline vec3 color_grid3(vec3 start, vec2 uv) {
    if (_grid_disable == 1) return start;
    uv = fract(uv * PE_L(0.5)) - vec2(PE_L(0.5), PE_L(0.5));
    real dist = max(abs(uv.x), abs(uv.y)) * PE_L(2.);
    if (dist > PE_L(0.985)) {
        return start * PE_L(0.4);
    } else if (dist < PE_L(0.94)) {
        return start;
    } else {
        if (uv.x > uv.y) {
            return start * PE_L(0.7);
        } else {
            return start * PE_L(1.2);
        }
    }
}

// library.glsl:286-288
static inline vec3 color_add_weighted(vec3 a, vec3 b, real coef) { return a * (PE_L(1.0) - coef) + b * coef; }

// library.glsl:297-301
struct MaterialProcessing {
    bool is_final;
    vec3 mul_to_color;
    Ray new_ray;
};

// library.glsl:303-305
static inline MaterialProcessing material_empty() { return MaterialProcessing{true, vec3(real(0)), ray_none}; }
// library.glsl:308-310
static inline MaterialProcessing material_final(vec3 color) { return MaterialProcessing{true, color, ray_none}; }
// library.glsl:313-315
static inline MaterialProcessing material_next(vec3 mul_color, Ray new_ray) {
    return MaterialProcessing{false, mul_color, new_ray};
}

// library.glsl:318-335
static inline MaterialProcessing material_simple2(SurfaceIntersection hit, Ray r, vec3 color, real normal_coef, bool grid,
                                                  real grid_scale, real grid_coef, bool grid2, bool grid3) {
    color = color_add_weighted(color, color * color_normal(hit.n, r.d), normal_coef);
    if (grid) {
        if (grid3) {
            color = color_add_weighted(color, color_grid3(color, vec2(hit.u, hit.v) * grid_scale), grid_coef);
        } else if (grid2) {
            color = color_add_weighted(color, color_grid2(color, vec2(hit.u, hit.v) * grid_scale), grid_coef);
        } else {
            color = color_add_weighted(color, color_grid(color, vec2(hit.u, hit.v) * grid_scale), grid_coef);
        }
    }
    return material_final(color);
}

// library.glsl:338-344
static inline MaterialProcessing material_simple(SurfaceIntersection hit, Ray r, vec3 color, real normal_coef, bool grid,
                                                 real grid_scale, real grid_coef) {
    return material_simple2(hit, r, color, normal_coef, grid, grid_scale, grid_coef, false, false);
}

// library.glsl:347-354
static inline MaterialProcessing material_reflect(SurfaceIntersection hit, Ray r, vec3 add_to_color) {
    r.d = vec4(my_reflect(vec3(r.d), hit.n), real(0));
    r.o += r.d * _offset_after_material;
    return material_next(add_to_color, r);
}

// library.glsl:357-364
static inline MaterialProcessing material_refract(SurfaceIntersection hit, Ray r, vec3 add_to_color, real refractive_index) {
    r.d = vec4(my_refract(vec3(r.d), hit.n, refractive_index), real(0));
    r.o += r.d * _offset_after_material;
    return material_next(add_to_color, r);
}

// library.glsl:366-371
static inline MaterialProcessing material_teleport_transformed(Ray r, vec3 n) {
    (void)n;
    r.o += r.d * _offset_after_material;
    r = normalize_ray(r);
    return material_next(vec3(real(1)), r);
}

// library.glsl:374-379
static inline MaterialProcessing material_teleport(SurfaceIntersection hit, Ray r, const mat4& teleport_matrix) {
    return material_teleport_transformed(transform(teleport_matrix, r), hit.n);
}

// library.glsl:381-384
static inline MaterialProcessing material_change_subspace(Ray r) {
    r.in_subspace = !r.in_subspace;
    return material_next(vec3(real(1)), r);
}

// library.glsl:387-398
#define CUSTOM_MATERIAL -1
#define NOT_INSIDE 0
#define TELEPORT 1
#define TELEPORT_SUBSPACE 2
#define DEBUG_RED 3
#define DEBUG_GREEN 4
#define DEBUG_BLUE 5
#define USER_MATERIAL_OFFSET 10

// library.glsl:405-409
struct SceneIntersection {
    int material;
    SurfaceIntersection hit;
    bool in_subspace;
};

// library.glsl:411
static const SceneIntersection scene_intersection_none = SceneIntersection{0, intersection_none, false};

// library.glsl:413-423
static inline bool nearer(const SurfaceIntersection& result, const SurfaceIntersection& current) {
    return current.hit && (current.t > real(0)) && (!result.hit || (result.hit && current.t < result.t));
}
static inline bool nearer(const SceneIntersection& result, const SurfaceIntersection& current) {
    return nearer(result.hit, current);
}
static inline bool nearer(const SceneIntersection& result, const SceneIntersection& current) {
    return nearer(result, current.hit);
}

// library.glsl:426-431
static inline vec3 cap_normal(vec3 pos, vec3 a, vec3 b, real radius) {
    vec3 ba = b - a;
    vec3 pa = pos - a;
    real h = clamp(dot(pa, ba) / dot(ba, ba), PE_L(0.0), PE_L(1.0));
    return (pa - h * ba) / radius;
}

// library.glsl:434-470
static inline SurfaceIntersection cap(Ray r, vec3 pa, vec3 pb, real radius) {
    vec3 ro = vec3(r.o);
    vec3 rd = vec3(r.d);
    vec3 ba = pb - pa;
    vec3 oa = ro - pa;

    real baba = dot(ba, ba);
    real bard = dot(ba, rd);
    real baoa = dot(ba, oa);
    real rdoa = dot(rd, oa);
    real oaoa = dot(oa, oa);

    real a = baba - bard * bard;
    real b = baba * rdoa - baoa * bard;
    real c = baba * oaoa - baoa * baoa - radius * radius * baba;
    real h = b * b - a * c;
    if (h >= PE_L(0.0)) {
        real t = (-b - sqrt(h)) / a;
        real y = baoa + t * bard;
        if (y > PE_L(0.0) && y < baba) {
            vec3 pos = ro + rd * t;
            return SurfaceIntersection{true, t, real(0), real(0), cap_normal(pos, pa, pb, radius)};
        }
        vec3 oc = (y <= PE_L(0.0)) ? oa : ro - pb;
        b = dot(rd, oc);
        c = dot(oc, oc) - radius * radius;
        h = b * b - c;
        if (h > PE_L(0.0)) {
            t = -b - sqrt(h);
            vec3 pos = ro + rd * t;
            return SurfaceIntersection{true, t, real(0), real(0), cap_normal(pos, pa, pb, radius)};
        }
    }
    return intersection_none;
}

// library.glsl:473-504
static inline SurfaceIntersection cylinder(Ray r, vec3 pa, vec3 pb, real ra) {
    vec3 ro = vec3(r.o);
    vec3 rd = vec3(r.d);
    vec3 ba = pb - pa;
    vec3 oc = ro - pa;

    real baba = dot(ba, ba);
    real bard = dot(ba, rd);
    real baoc = dot(ba, oc);

    real k2 = baba - bard * bard;
    real k1 = baba * dot(oc, rd) - baoc * bard;
    real k0 = baba * dot(oc, oc) - baoc * baoc - ra * ra * baba;

    real h = k1 * k1 - k2 * k0;
    if (h < PE_L(0.0)) return intersection_none;
    h = sqrt(h);

    real t = (-k1 - h) / k2;
    real y = baoc + t * bard;
    if (y > PE_L(0.0) && y < baba)
        return SurfaceIntersection{true, t, real(0), real(0), (oc + t * rd - ba * y / baba) / ra};

    t = (-k1 + h) / k2;
    y = baoc + t * bard;
    if (y > PE_L(0.0) && y < baba)
        return SurfaceIntersection{true, t, real(0), real(0), (oc + t * rd - ba * y / baba) / ra};

    return intersection_none;
}

// library.glsl:507-525
static inline SurfaceIntersection triangle(Ray r, vec3 v0, vec3 v1, vec3 v2) {
    vec3 ro = vec3(r.o);
    vec3 rd = vec3(r.d);

    vec3 v1v0 = v1 - v0;
    vec3 v2v0 = v2 - v0;
    vec3 rov0 = ro - v0;

    vec3 n = cross(v1v0, v2v0);
    vec3 q = cross(rov0, rd);
    real d = PE_L(1.0) / dot(rd, n);
    real u = d * dot(-q, v2v0);
    real v = d * dot(q, v1v0);
    real t = d * dot(-n, rov0);

    if (u < PE_L(0.0) || v < PE_L(0.0) || (u + v) > PE_L(1.0)) return intersection_none;

    return SurfaceIntersection{true, t, u, v, normalize_normal(cross(v1 - v0, v2 - v0), vec3(r.d))};
}

// library.glsl:528-554
static inline SceneIntersection debug_intersect(Ray r) {
    vec3 pa = vec3(real(0));
    real radius = PE_L(0.03);

    SurfaceIntersection hit = intersection_none;
    SceneIntersection i = SceneIntersection{0, hit, false};

    hit = cap(r, pa, vec3(real(1), real(0), real(0)), radius);
    if (nearer(i, hit)) {
        i.material = DEBUG_RED;
        i.hit = hit;
    }
    hit = cap(r, pa, vec3(real(0), real(1), real(0)), radius);
    if (nearer(i, hit)) {
        i.material = DEBUG_GREEN;
        i.hit = hit;
    }
    hit = cap(r, pa, vec3(real(0), real(0), real(1)), radius);
    if (nearer(i, hit)) {
        i.material = DEBUG_BLUE;
        i.hit = hit;
    }
    return i;
}

// library.glsl:560-572
static inline SceneIntersection process_plane_intersection(SceneIntersection i, SurfaceIntersection hit, int inside) {
    if (inside == NOT_INSIDE) {
    } else if (inside == TELEPORT) {
    } else if (inside == TELEPORT_SUBSPACE) {
    } else {
        i.hit = hit;
        i.material = inside;
    }
    return i;
}

// library.glsl:574-589
static inline SceneIntersection process_portal_intersection(SceneIntersection i, SurfaceIntersection hit, int inside,
                                                            int teleport_material) {
    if (inside == NOT_INSIDE) {
    } else if (inside == TELEPORT) {
        i.hit = hit;
        i.material = teleport_material;
    } else if (inside == TELEPORT_SUBSPACE) {
        i.hit = hit;
        i.material = teleport_material;
        i.in_subspace = true;
    } else {
        i.hit = hit;
        i.material = inside;
    }
    return i;
}

// library.glsl:595-598
struct SceneIntersectionWithMaterial {
    SceneIntersection scene;
    MaterialProcessing material;
};

}  // namespace pe_oracle
