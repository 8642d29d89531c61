"""Markdown table of a tools/sweep.py result (BASELINE.json configs[4]).
usage: python tools/sweep_table.py gpurun_out/r02_sweep.json > profiles/r02_sweep_table.md"""
import json
import sys

d = json.load(open(sys.argv[1]))
rows = d["rows"]
peak = d.get("hbm_peak_GBps")
print("# BASELINE.json configs[4]: FPS + ball_query + group_points sweep (B=8 scenes, M=2048 centres, 1xB200)\n")
print("`python tools/sweep.py` (raw rows: `%s`). Room scenes of `scan2cap_b200/synthetic.py`, radius 0.2 m x sqrt(40000/N) (constant"
      % sys.argv[1].replace("gpurun_out", "profiles"))
print("expected ball occupancy), L2 flushed before every timed call, median of 7.  `frac` = algorithmic bytes (SURVEY.md 8(d)) / time /")
print("%.0f GB/s (%s).  ref = the reference's lib/pointnet2 kernels (sm_100 build) on the same inputs.  The fused op is measured in the"
      % (peak, d.get("peak_source")))
print("product layout (channels-last rows `[x,y,z,0 | features | pad]`), grid build included.\n")
if d.get("note"):
    print(d["note"] + "\n")


def f(v, fmt):
    return "-" if v is None else fmt % v


print("## fused query+group (QueryAndGroup.forward in one entry point: grid build + query/gather)\n")
print("| N | C | nsample | ours ms | ref ms | speed-up | alg MB | GB/s | frac HBM | note |\n|---|---|---|---|---|---|---|---|---|---|")
for r in sorted((r for r in rows if r["op"] == "query_and_group"), key=lambda r: (r["N"], r["C"], r["ns"])):
    sp = (r["ref_ms"] / r["ours_ms"]) if r.get("ref_ms") else None
    note = ("latency-scale byte floor (%.1f us)" % r["byte_floor_us"]) if r.get("note") else ""
    print("| %d | %d | %d | %.3f | %s | %s | %.1f | %.0f | %.1f %% | %s |" % (
        r["N"], r["C"], r["ns"], r["ours_ms"], f(r.get("ref_ms"), "%.1f"), f(sp, "%.0fx"), r["alg_bytes"] / 1e6, r["GBps"],
        100 * r["frac_hbm"], note))
print("\n## ball_query alone (C irrelevant)\n")
print("| N | nsample | ours ms | ref ms | speed-up | mean ball fill |\n|---|---|---|---|---|---|")
for r in sorted((r for r in rows if r["op"] == "ball_query"), key=lambda r: (r["N"], r["ns"])):
    sp = (r["ref_ms"] / r["ours_ms"]) if r.get("ref_ms") else None
    print("| %d | %d | %.3f | %s | %s | %.2f |" % (r["N"], r["ns"], r["ours_ms"], f(r.get("ref_ms"), "%.1f"), f(sp, "%.0fx"),
                                               r.get("mean_ball_fill", float("nan"))))
gp = [r for r in rows if r["op"] == "group_points"]
if gp:
    print("\n## group_points alone (the `_ext` API on (B,C,N))\n")
    print("| N | C | nsample | ours ms | ref ms | speed-up | GB/s | frac HBM |\n|---|---|---|---|---|---|---|---|")
    for r in sorted(gp, key=lambda r: (r["N"], r["C"], r["ns"])):
        sp = (r["ref_ms"] / r["ours_ms"]) if r.get("ref_ms") else None
        print("| %d | %d | %d | %.3f | %s | %s | %.0f | %.1f %% |" % (r["N"], r["C"], r["ns"], r["ours_ms"], f(r.get("ref_ms"), "%.3f"),
                                                                    f(sp, "%.1fx"), r["GBps"], 100 * r["frac_hbm"]))
print("\n## furthest point sampling (latency bound: a serial chain of M arg-max picks)\n")
print("| N | ours ms | ref ms | speed-up | us / pick |\n|---|---|---|---|---|")
for r in sorted((r for r in rows if r["op"] == "fps"), key=lambda r: r["N"]):
    sp = (r["ref_ms"] / r["ours_ms"]) if r.get("ref_ms") else None
    print("| %d | %.3f | %s | %s | %.2f |" % (r["N"], r["ours_ms"], f(r.get("ref_ms"), "%.1f"), f(sp, "%.1fx"), r["us_per_pick"]))
