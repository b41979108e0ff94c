"""Random, well-typed GLSL expressions for the differential test of the two GLSL layers (tests/test_program_on_host.py).

Every expression is built from ray / hit quantities and literals with the operators, swizzles (narrowing, permuting and
widening), constructors and built-ins of GLSL ES 3.00 sections 5 and 8 in all their genType forms; `snippet(seed)` packs a
number of them into one material body whose colour depends on every finite one."""
import random

F1 = ["sin","cos","tan","asin","acos","atan","exp","log","exp2","log2","sqrt","inversesqrt","abs","sign","floor","ceil","fract","radians","degrees","trunc","round","roundEven","sinh","cosh","tanh"]
F2 = ["min","max","mod","pow","step","atan"]
N = {"f":1,"v2":2,"v3":3,"v4":4}
CT = {"f":"float","v2":"vec2","v3":"vec3","v4":"vec4"}

def gen(rng, t, d):
    k = rng.random()
    if d <= 0 or k < 0.18:
        if t == "f":
            return rng.choice(["hit.u","hit.v","hit.t","r.d.x","r.d.y","hit.n.z","0.5","2.","1.5e-1",".25","3.","-1.","0.","1e3","1e-4", "float(3)", "r.o.y"])
        if t == "v2": return rng.choice(["vec2(hit.u, hit.v)","hit.n.xy","r.d.zx","vec2(0.5)","r.o.xz"])
        if t == "v3": return rng.choice(["hit.n","r.d.xyz","r.o.xyz","vec3(hit.u, hit.v, hit.t)","vec3(0.3)","hit.n.zxy","r.d.yyx"])
        return rng.choice(["r.d","r.o","vec4(hit.n, 1.)","vec4(hit.u)","r.d.wzyx","vec4(r.o.xy, hit.n.yz)"])
    if k < 0.40:
        op = rng.choice(["+","-","*","/"])
        if t != "f" and rng.random() < 0.35:
            if rng.random() < 0.5: return f"({gen(rng,t,d-1)} {op} {gen(rng,'f',d-1)})"
            return f"({gen(rng,'f',d-1)} {op} {gen(rng,t,d-1)})"
        return f"({gen(rng,t,d-1)} {op} {gen(rng,t,d-1)})"
    if k < 0.45: return f"(-({gen(rng,t,d-1)}))"        # parenthesised: `--1.` would be a decrement
    if k < 0.62: return f"{rng.choice(F1)}({gen(rng,t,d-1)})"
    if k < 0.72:
        f = rng.choice(F2)
        if t != "f" and f in ("min","max","mod") and rng.random() < 0.4: return f"{f}({gen(rng,t,d-1)}, {gen(rng,'f',d-1)})"
        if t != "f" and f == "step" and rng.random() < 0.4: return f"step({gen(rng,'f',d-1)}, {gen(rng,t,d-1)})"
        return f"{f}({gen(rng,t,d-1)}, {gen(rng,t,d-1)})"
    if k < 0.78:
        c = rng.random()
        if c < 0.4: return f"clamp({gen(rng,t,d-1)}, {gen(rng,'f',d-1)}, {gen(rng,'f',d-1)})"
        if c < 0.7: return f"mix({gen(rng,t,d-1)}, {gen(rng,t,d-1)}, {gen(rng,'f',d-1)})"
        if c < 0.85 or t == "f": return f"smoothstep({gen(rng,'f',d-1)}, {gen(rng,'f',d-1)}, {gen(rng,t,d-1)})"
        return f"mix({gen(rng,t,d-1)}, {gen(rng,t,d-1)}, {gen(rng,t,d-1)})"
    if t == "f":
        vt = rng.choice(["v2","v3","v4"])
        c = rng.random()
        if c < 0.3: return f"length({gen(rng,vt,d-1)})"
        if c < 0.6: return f"dot({gen(rng,vt,d-1)}, {gen(rng,vt,d-1)})"
        if c < 0.75: return f"distance({gen(rng,vt,d-1)}, {gen(rng,vt,d-1)})"
        return f"{gen(rng,vt,d-1)}.{rng.choice('xy' if vt=='v2' else ('xyz' if vt=='v3' else 'xyzw'))}"
    c = rng.random()
    if c < 0.2: return f"normalize({gen(rng,t,d-1)})"
    if c < 0.3 and t == "v3": return f"cross({gen(rng,'v3',d-1)}, {gen(rng,'v3',d-1)})"
    if c < 0.42: return f"reflect({gen(rng,t,d-1)}, {gen(rng,t,d-1)})"
    if c < 0.5: return f"refract({gen(rng,t,d-1)}, {gen(rng,t,d-1)}, {gen(rng,'f',d-1)})"
    if c < 0.56: return f"faceforward({gen(rng,t,d-1)}, {gen(rng,t,d-1)}, {gen(rng,t,d-1)})"
    if c < 0.75:   # constructor from parts
        if t == "v2": return f"vec2({gen(rng,'f',d-1)}, {gen(rng,'f',d-1)})"
        if t == "v3": return rng.choice([f"vec3({gen(rng,'v2',d-1)}, {gen(rng,'f',d-1)})", f"vec3({gen(rng,'f',d-1)}, {gen(rng,'v2',d-1)})", f"vec3({gen(rng,'f',d-1)})"])
        return rng.choice([f"vec4({gen(rng,'v3',d-1)}, {gen(rng,'f',d-1)})", f"vec4({gen(rng,'v2',d-1)}, {gen(rng,'v2',d-1)})", f"vec4({gen(rng,'f',d-1)}, {gen(rng,'v3',d-1)})"])
    # swizzle from another vector
    src = rng.choice(["v2","v3","v4"])
    comps = "xy" if src=="v2" else ("xyz" if src=="v3" else "xyzw")
    return f"{gen(rng,src,d-1)}." + "".join(rng.choice(comps) for _ in range(N[t]))

def snippet(seed, n_expr=24, depth=4):
    rng = random.Random(seed)
    lines = ["vec3 acc = vec3(0.);", "float cnt = 0.;"]
    for i in range(n_expr):
        t = rng.choice(["f","v2","v3","v4"])
        e = gen(rng, t, depth)
        lines.append(f"{CT[t]} e{i} = {e};")
        red = {"f": f"vec3(e{i})", "v2": f"vec3(e{i}, e{i}.x)", "v3": f"e{i}", "v4": f"e{i}.xyz + vec3(e{i}.w)"}[t]
        # keep only finite, bounded contributions so that one Inf/NaN does not blank the whole pixel
        lines.append(f"{{ vec3 q = {red}; if (!(abs(q.x) > 1e6) && !(abs(q.y) > 1e6) && !(abs(q.z) > 1e6) && q.x == q.x && q.y == q.y && q.z == q.z) {{ acc += q; cnt += 1.; }} }}")
    lines.append("vec3 c = fract(abs(acc) * 0.37 + vec3(cnt * 0.01));")
    lines.append("return material_simple(hit, r, c, 5e-1, false, 4e0, 3e-1);")
    return "\n".join(lines)
