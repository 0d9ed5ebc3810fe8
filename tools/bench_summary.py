#!/usr/bin/env python3
"""Summarise bench.py JSON lines (one per log file) into a markdown table for profiles/.

usage: python tools/bench_summary.py title log1 [log2 ...] > profiles/rNN_summary.md
"""
import json
import sys


def line_of(path):
    last = None
    with open(path, errors="replace") as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith("{") and '"metric"' in ln:
                try:
                    last = json.loads(ln)
                except Exception:
                    pass
    return last


def main():
    title = sys.argv[1]
    print("# %s\n" % title)
    for path in sys.argv[2:]:
        j = line_of(path)
        print("## `%s`\n" % path)
        if j is None:
            print("no JSON line found\n")
            continue
        cfg = j.get("config", {})
        print("* workload: %s" % cfg.get("workload"))
        print("* lanes: %s; sharding: %s; GPUs: %s" % (cfg.get("lanes"), cfg.get("sharding"), j.get("n_gpus")))
        print("* **value %.0f %s**, %.1f ms / step (%d steps, %d warm-up)" % (j["value"], j["unit"], j["ms_per_step"], j["steps"], j["warmup"]))
        ph = j.get("phases_ms") or {}
        print("* phases (ms): " + ", ".join("%s %.2f" % (k.replace("ms_", ""), v) for k, v in ph.items()))
        for key in ("roofline", "roofline_k1", "roofline_step"):
            r = j.get(key)
            if r:
                print("* %s: %s — %.1f GB/s of %.1f (%s) = **%.4f**; %.2f ms / launch, algorithmic %.2f GB%s" % (
                    key, r.get("kernel"), r["achieved"], r["peak"], r.get("peak_source", ""), r["frac"], r["ms_per_launch"],
                    r["algorithmic_bytes_per_launch"] / 1e9, ("; DRAM traffic %.1f GB" % (r["traffic"] / 1e9)) if r.get("traffic") else ""))
        e = j.get("e2e")
        if e:
            print("* e2e: %.0f %s over %s steps, %d streams, H2D %.1f GB / step, same frames as resident: %s" % (
                e["value"], e["unit"], e.get("steps"), e.get("streams", 0), e["h2d_bytes_per_step"] / 1e9, e.get("same_frames_as_resident")))
            s16 = e.get("s16")
            if s16 and "value" in s16:
                print("* e2e int16 IQ: %.0f %s, H2D %.1f GB / step, %d frames" % (s16["value"], s16["unit"], s16["h2d_bytes_per_step"] / 1e9, s16.get("frames_per_step", 0)))
        c = j.get("cpu_baseline")
        if c:
            print("* cpu_baseline: %.1f %s on %d threads (%s): %s" % (c["value"], c["unit"], c["cores"], c["kind"], c["sample"]))
        fp = j.get("full_parity")
        if fp:
            print("* full_parity: %s" % json.dumps(fp))
        fc = j.get("full_size_check")
        if fc:
            print("* full_size_check: %s" % json.dumps(fc))
        ws = j.get("wav_set")
        if ws:
            print("* wav_set: %s" % json.dumps(ws))
        d = j.get("decode")
        if d:
            print("* decode: %s" % json.dumps(d))
        ck = j.get("clocks")
        if ck:
            print("* clocks: %s" % json.dumps(ck))
        print("* frames_digest %s, gpu_launches %s" % (j.get("frames_digest"), j.get("gpu_launches")))
        print()


if __name__ == "__main__":
    main()
