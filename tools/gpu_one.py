import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nfcutil as U
import nfc_laboratory_b200 as N
name = sys.argv[1] if len(sys.argv) > 1 else "test_NFC-A_106kbps_002"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mag, rate, _ = U.fixture_wav(name)
if n: mag = mag[:n]
d = N.NfcDecoder()
fr = d.decode_batch(mag[None], N.SIG_MAG_F32, rate)
print(len(fr), d.stats()["ms_lanes"])
d.close()
