"""streaming probe: pushes one fixture through nextFrames in 65536-sample buffers and prints progress (debug aid)"""
import faulthandler, os, sys
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nfcutil as U
import nfc_laboratory_b200 as N
name = sys.argv[1] if len(sys.argv) > 1 else "test_NFC-A_106kbps_001"
mag, rate, _ = U.fixture_wav(name)
d = N.NfcDecoder()
got = []
for pos in range(0, mag.size, 65536):
    fr = d.nextFrames(mag[pos:pos + 65536], rate)
    print("push", pos, len(fr), flush=True)
    got += fr
got += d.nextFrames(None)
print("flush", len(got), flush=True)
d.close()
print("closed", flush=True)
