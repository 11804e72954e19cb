"""Warp-stall sampling breakdown of one kernel from an .ncu-rep captured with `--set full --import-source on`.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep mp_headtile [--top 30] [--landmarks]

Prints the share of samples per stall reason, the hottest SASS instructions, and (with --landmarks) the cumulative
share at every barrier / fence / global access, which is how profiles/r1_mp_stall_breakdown.md splits the kernel
into regions.  Needs only `ncu` on the PATH (no GPU)."""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel_regex")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--landmarks", action="store_true")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel_regex}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][1] if start else "")
    hdr, data = rows[start], []
    for r in rows[start + 1:]:
        if r and r[0] == "Kernel Name":
            break  # first matching launch only
        if len(r) == len(hdr):
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: sum(f(r, s) for r in data) for s in stalls}
    total = sum(tot.values()) or 1.0
    print(f"samples {total:.0f}, SASS instructions {len(data)}")
    for s, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v:
            print(f"  {s:26s} {v:8.0f} {100 * v / total:5.1f} %")
    print("hottest instructions")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:a.top]:
        main_stall = max(stalls, key=lambda s: f(r, s))
        print(f"  {r[ix['Address']][-5:]} {100 * f(r, '# Samples') / total:5.1f} %  executed {f(r, 'Instructions Executed'):9.0f}  "
              f"{main_stall:22s} {r[ix['Source']][:72]}")
    if a.landmarks:
        print("cumulative share at landmarks (address order)")
        acc = 0.0
        for r in data:
            acc += f(r, "# Samples")
            src = r[ix["Source"]]
            if any(k in src for k in ("SYNCS", "FENCE", "MEMBAR", "BAR.", "LDG", "STG", "EXIT", "UBLKCP", "UTMA")) and \
                    f(r, "Instructions Executed") > 0:
                print(f"  {r[ix['Address']][-5:]} {100 * acc / total:5.1f} %  executed {f(r, 'Instructions Executed'):9.0f}  {src[:72]}")


if __name__ == "__main__":
    main()
