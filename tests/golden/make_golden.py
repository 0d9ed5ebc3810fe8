#!/usr/bin/env python3
"""Generate tests/golden/ from the reference checkout (run in the build container only; /root/reference is absent
on the GPU box, so the fixtures are committed).

For every regression capture wav/test_*.wav of the reference:
  * <name>.wav.xz  -- the capture, xz-compressed (mono 16-bit PCM @ 10 MS/s)
  * <name>.json    -- the reference's OWN golden frame list, copied verbatim (what test-sdr compares against)
  * <name>.ref.json-- every frame the compiled reference (oracle/_ref/libnfcref.so) emits for the capture, carrier
                      on/off frames included (those are filtered out of <name>.json, test-sdr main.cpp:171-174)

Usage: python tests/golden/make_golden.py [/root/reference]
"""
import json
import lzma
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import nfcutil  # noqa: E402


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    wav = os.path.join(ref, "wav")
    for fn in sorted(os.listdir(wav)):
        if not fn.endswith(".wav"):
            continue
        name = fn[:-4]
        with open(os.path.join(wav, fn), "rb") as f:
            raw = f.read()
        dst = os.path.join(HERE, name + ".wav.xz")
        if not os.path.exists(dst):
            with lzma.open(dst, "wb", preset=9) as f:
                f.write(raw)
        shutil.copyfile(os.path.join(wav, name + ".json"), os.path.join(HERE, name + ".json"))
        mag, rate, _ = nfcutil.read_wav(os.path.join(wav, fn))
        frames = nfcutil.ref_decode(mag, rate)
        with open(os.path.join(HERE, name + ".ref.json"), "w") as f:
            json.dump({"sampleRate": rate, "samples": int(mag.size),
                       "frames": [[t, ty, fl, ph, r, s, e, d.hex()] for (t, ty, fl, ph, r, s, e, d) in frames]}, f, indent=0)
        print("%-40s %8d samples %4d frames" % (name, mag.size, len(frames)))


if __name__ == "__main__":
    main()
