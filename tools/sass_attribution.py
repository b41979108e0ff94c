#!/usr/bin/env python
"""Where the executed instructions of pe_render_kernel go, by SOURCE CONTEXT: joins the per-instruction execution counts of an
`ncu --set full` report (SASS page) with the line table (+ inline chains) of the same cubin, rebuilt here from the scene.

    python tools/sass_attribution.py gpurun_out/r02e_ncu_portal_in_portal.ncu-rep portal_in_portal [--options '{"canon_rays":1}']
                                     [--top 40] > profiles/r02e_attribution_portal_in_portal.txt

The generated program carries `#line` directives, so a line is "<generated>" (kernel + generator sections), "scene_program.cu"
(the device headers) or the scene element that owns a GLSL snippet.  Three tables: (1) share per generator section -- which
call of bounce_body the instruction was inlined under (scene_intersect / the intersection materials / material_process / ray
set-up); (2) share per innermost source line; (3) the opcodes of the hottest section.  The cubin compiled here must be the
one that ran (same tree, same NVRTC: the library loads the toolkit's by path); the tool checks opcode by opcode and refuses
otherwise.  Needs ncu, nvdisasm (reads reports only; no GPU)."""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ncu_counts(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    while rows and "Source" not in rows[0]:
        rows = rows[1:]
    hdr = rows[0]
    i_adr, i_src, i_cnt = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    body = [r for r in rows[1:] if len(r) > i_cnt]
    base = int(body[0][i_adr], 16)
    return {int(r[i_adr], 16) - base: (r[i_src].strip(), int(float(r[i_cnt]))) for r in body}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("scene")
    ap.add_argument("--options", default="{}")
    ap.add_argument("--persistent", type=int, default=0)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--ops", default="", help="comma-separated opcodes: add a table of source lines by executed instructions of these opcodes only (e.g. BSSY,BSYNC,BRA)")
    args = ap.parse_args()
    from portal_b200.renderer import SceneRenderer, load_scene_ir
    ir = load_scene_ir(os.path.join(ROOT, "tests", "golden", "scenes", f"{args.scene}.scene.json"))
    r = SceneRenderer(ir, device=-1, persistent=bool(args.persistent), options=json.loads(args.options))
    src = r.source()
    cubin = f"/tmp/_attr_{os.getpid()}.cubin"
    open(cubin, "wb").write(r.cubin())
    nvdi = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    os.remove(cubin)
    counts = ncu_counts(args.report)

    text = collections.defaultdict(dict)
    f, ln = "scene_program.cu", 0
    for raw in src.split("\n"):
        ln += 1
        m = re.match(r'#line (\d+) "([^"]+)"', raw)
        if m:
            f, ln = m.group(2), int(m.group(1)) - 1
            continue
        text[f][ln] = raw.strip()

    seq, stack, pending, intext = [], [], [], False
    for line in nvdi.split("\n"):
        if line.startswith("//---"):
            intext = ".text.pe_render_kernel" in line
        if not intext:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            pending.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                stack, pending = pending, []
            seq.append((int(m.group(1), 16), m.group(2).strip(), tuple(stack)))

    def opcode(s):
        p = s.split()
        return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
    bad = sum(1 for a, ins, _ in seq if a not in counts or opcode(counts[a][0]) != opcode(ins))
    if bad:
        sys.exit(f"the cubin compiled here is not the one in the report ({bad} of {len(seq)} opcodes differ): same tree, options and NVRTC?")
    total = sum(counts[a][1] for a, _, _ in seq)

    def section(st):
        chain = " | ".join(text.get(f_, {}).get(l_, "") for f_, l_ in reversed(st))
        pre = "non-canonical copy: " if "o.done = bounce_body" in chain else ""
        if "= scene_intersect_material_process(r)" in chain:
            return pre + "intersection materials (scene_intersect_material_process)"
        if "= scene_intersect(r)" in chain:
            return pre + "objects (scene_intersect)"
        if "m = material_process(" in chain:
            return pre + "materials (material_process)"
        if "primary_ray(" in chain:
            return "primary ray (pixel -> ray, camera)"
        if "store_pixel(" in chain:
            return "store (sqrt, float4 / RGBA8)"
        if pre:
            return pre + "rest of the bounce"
        if "bounce_once(" in chain or "bounce_body(" in chain:
            return "rest of the bounce (advance, darken, loop control)"
        # code of a function the compiler did not inline has no chain up to the kernel: name it by the scene element / header it is in
        owner = st[-1][0] if st else "?"
        if owner not in ("<generated>", "?"):
            return f"out-of-line function in {owner}" if owner != "scene_program.cu" else "out-of-line device-header function (called from a snippet)"
        return "kernel prologue / index arithmetic"

    by_sec, by_line, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for a, ins, st in seq:
        c = counts[a][1]
        s_ = section(st)
        by_sec[s_] += c
        by_line[st[0] if st else ("?", 0)] += c
        ops[s_][opcode(ins)] += c
    px = float(counts[min(counts)][1]) or 1.0     # the first instruction runs once per warp
    print(f"# {args.report}: {total} warp-instructions, {len(seq)} static, {total / px:.0f} per warp (= per thread)")
    print("\n## share per generator section")
    for s_, c in by_sec.most_common():
        print(f"{100 * c / total:6.2f}%  {c / px:8.1f} instr/thread  {s_}")
    print("\n## share per innermost source line")
    for (f_, l_), c in by_line.most_common(args.top):
        print(f"{100 * c / total:6.2f}%  {f_[:30]:30s}:{l_:<5d} {text.get(f_, {}).get(l_, '')[:120]}")
    if args.ops:
        want = set(args.ops.split(","))
        sel = collections.Counter()
        for a, ins, st in seq:
            if opcode(ins) in want:
                sel[st[0] if st else ("?", 0)] += counts[a][1]
        tot_s = sum(sel.values())
        print(f"\n## source lines by {args.ops} only ({100 * tot_s / total:.2f}% of all executed instructions)")
        for (f_, l_), c in sel.most_common(args.top):
            print(f"{100 * c / total:6.2f}%  {f_[:30]:30s}:{l_:<5d} {text.get(f_, {}).get(l_, '')[:120]}")
    hot = by_sec.most_common(1)[0][0]
    print(f"\n## opcodes of `{hot}`")
    tot_h = sum(ops[hot].values())
    for op, c in ops[hot].most_common(16):
        print(f"{100 * c / tot_h:6.2f}%  {op}")


if __name__ == "__main__":
    main()
