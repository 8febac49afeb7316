"""Hot source lines of one kernel from an `ncu --set full --import-source on` report (compiled with -lineinfo).
usage: python profiles/ncu_source_lines.py report.ncu-rep <kernel regex> [launch-skip] [top N]"""
import csv
import io
import subprocess
import sys


def main(path, kernel, skip="0", top="25"):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fpath, lines, hdr = "", [], None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            print(r[1][:140])
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr[:10], r[:10]))
            try:
                lines.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), fpath, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    ti, ts = sum(l[0] for l in lines), sum(l[1] for l in lines)
    print("total warp instructions %d, samples %d" % (ti, ts))
    for inst, samp, f, ln, src in sorted(lines, key=lambda l: -l[1])[: int(top)]:
        print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100.0 * samp / max(ts, 1), 100.0 * inst / max(ti, 1), f, ln, src))


if __name__ == "__main__":
    main(*sys.argv[1:])
