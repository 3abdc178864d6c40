"""Full text extract of an .ncu-rep for profiles/: EVERY metric of the raw page for one launch (name [unit] = value) and the
hottest SASS / source lines of the source page (warp-stall samples), so the committed evidence is the whole capture in text form
rather than a 20-line digest.

    python scripts/ncu_full_extract.py rep.ncu-rep out.txt [launch_index=last] [top_lines=40]
"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40


def page(name, extra=()):
    r = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


rows = page("raw")
hdr, units, launches = rows[0], rows[1], rows[2:]
vals = launches[which]
lines = [f"# {rep}: launch {which if which >= 0 else len(launches) + which} of {len(launches)} captured; every metric of `ncu --page raw`",
         "# (`ncu --set full --clock-control none --import-source on`; times under the profiler are NOT bench values)", ""]
for h, u, v in zip(hdr, units, vals):
    if v != "":
        lines.append(f"{h} [{u}] = {v}")

# source page: per-instruction sampling data of the same launch
src = page("source", ("--print-source", "sass"))
hi = next((i for i, r in enumerate(src) if r and r[0] == "Address"), None)
if hi is not None:
    h = src[hi]
    col = {n: i for i, n in enumerate(h)}
    samp, srcc = "Warp Stall Sampling (All Samples)", "Source"
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    body = []
    for r in src[hi + 1:]:
        if len(r) != len(h):
            continue
        try:
            s = float(r[col[samp]] or 0)
        except ValueError:
            continue
        why = max(stalls, key=lambda n: float(r[col[n]] or 0)) if s > 0 else ""
        body.append((s, r[col[srcc]].strip(), why, r[col["Instructions Executed"]]))
    tot = sum(b[0] for b in body) or 1.0
    lines += ["", f"# source page: {len(body)} SASS instructions, {tot:.0f} warp-stall samples; the {top} hottest (samples, share, dominant stall reason, warp-level executions, SASS)"]
    for s, t, why, n in sorted(body, key=lambda b: -b[0])[:top]:
        lines.append(f"{s:9.0f}  {100 * s / tot:5.1f}%  {why:<22} {n:>10}  {t}")
    per, cnt = {}, {}
    for s, t, why, n in body:
        w = t.split()
        op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?")).split(".")[0]
        per[op] = per.get(op, 0.0) + s
        cnt[op] = cnt.get(op, 0) + int(n or 0)
    lines += ["", "# by opcode: samples, share, warp-level executions (top 24 by samples)"]
    for op, s in sorted(per.items(), key=lambda kv: -kv[1])[:24]:
        lines.append(f"{s:9.0f}  {100 * s / tot:5.1f}%  {cnt[op]:>12}  {op}")
    tot_r = {n: sum(float(r[col[n]] or 0) for r in src[hi + 1:] if len(r) == len(h)) for n in stalls}
    lines += ["", "# stall reasons over the whole kernel (samples)"]
    for n, v in sorted(tot_r.items(), key=lambda kv: -kv[1]):
        if v > 0:
            lines.append(f"{v:9.0f}  {100 * v / tot:5.1f}%  {n}")
open(out, "w").write("\n".join(lines) + "\n")
print(f"{out}: {len(lines)} lines")
