"""Per-source-line instruction counts of one profiled kernel, here on the CPU box.

    python profiles/sass_by_line.py gpurun_out/x.ncu-rep <mangled-or-substring of kernel name> [launch index] [top N]

`ncu --page source --csv` lists the executed-instruction count per SASS instruction but not its source line;
`nvdisasm --print-line-info-inline` of the cubin inside libsoccer_b200.so (built with -lineinfo) lists the line of
every SASS instruction.  The two are joined on the instruction offset and summed per (file, line), innermost
inlined frame.  The library must be the build that was profiled.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gym_soccer_littman94_b200", "libsoccer_b200.so")


def sass_lines(kernel_sub):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(td, cubin)], check=True,
                             capture_output=True, text=True).stdout
    out, cur, loc, pending = {}, None, None, None
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur = m.group(1) if kernel_sub in m.group(1) else None
            loc = pending = None
            continue
        if cur is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if pending is None:                      # the first frame listed is the innermost
                pending = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending is not None:
                loc, pending = pending, None
            out.setdefault(cur, {})[int(m.group(1), 16)] = (loc, m.group(2).strip())
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    # one block per launch: "Kernel Name",... then a header row then SASS rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
    blk = blocks[launch]
    hdr = blk["hdr"]
    ia, ii, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    base = int(blk["rows"][0][ia], 16)
    fmap = sass_lines(ksub)
    assert len(fmap) == 1, f"kernel substring matches {list(fmap)}"
    lines = next(iter(fmap.values()))
    per, tot, tot_t, unmatched = {}, 0, 0, 0
    for r in blk["rows"]:
        off, n, nt = int(r[ia], 16) - base, int(r[ii]), int(r[it])
        tot += n
        tot_t += nt
        loc = lines.get(off, (None, ""))[0]
        if loc is None:
            unmatched += n
        per.setdefault(loc, [0, 0])
        per[loc][0] += n
        per[loc][1] += nt
    print(f"{blk['name'][:100]}\n{tot} warp instructions, {tot_t} thread instructions, unmatched {unmatched}")
    srcs = {}
    for loc, (n, nt) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            path = os.path.join(ROOT, "gym_soccer_littman94_b200", "csrc", loc[0])
            if os.path.exists(path):
                srcs.setdefault(path, open(path).read().splitlines())
                text = srcs[path][loc[1] - 1].strip()[:100] if loc[1] - 1 < len(srcs[path]) else ""
        print(f"{100.0 * n / tot:5.1f}%  {n:>11d}  {str(loc):32s} {text}")


if __name__ == "__main__":
    main()
