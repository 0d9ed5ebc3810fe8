"""Frame lists in the reference's output formats (nfc_laboratory_b200/export.py): the regression tool's golden JSON,
the TRZ container (checked with the reference's own Python reader where the reference tree is present) and nfc-rx lines."""
import json
import os
import sys
import tarfile

import pytest

import nfcutil as U
from test_golden_oracle import committed_ref

NAMES = U.fixture_names()
REF_TOOLS = "/root/reference/tools"


@pytest.mark.parametrize("name", NAMES)
def test_golden_json_is_reproduced_from_the_frame_list(name, tmp_path):
    """frames (recorded output of the compiled reference) -> write_frames_json == the reference's golden file, value for
    value: sample indices, the double time stamps, flags, payload text"""
    from nfc_laboratory_b200 import export as X
    frames = committed_ref(name)[0]
    p = tmp_path / "out.json"
    X.write_frames_json(p, frames, 10_000_000)
    with open(p) as f:
        ours = json.load(f)
    with open(os.path.join(U.GOLDEN, name + ".json")) as f:
        golden = json.load(f)
    assert ours == golden


def test_trz_roundtrip_and_container(tmp_path):
    from nfc_laboratory_b200 import export as X
    frames = committed_ref("test_NFC-A_424kbps_002")[0]
    p = tmp_path / "trace.trz"
    X.write_trz(p, frames, 10_000_000, stream_time=1_700_000_000)
    assert X.read_trz(p) == frames
    with tarfile.open(p, "r:gz") as tar:
        assert tar.getnames() == ["frame.json"]
        doc = json.load(tar.extractfile("frame.json"))
    first = doc["frames"][2]
    assert set(first) >= {"sampleStart", "sampleEnd", "sampleRate", "timeStart", "timeEnd", "techType", "frameType", "frameRate", "frameFlags", "framePhase", "dateTime"}
    assert first["dateTime"] == 1_700_000_000 + first["timeStart"]
    carrier = [e for e in doc["frames"] if e["frameType"] in (0x100, 0x101)]
    assert carrier and all("frameData" not in e for e in carrier)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_TOOLS, "py_nfclab")), reason="reference tree not present")
def test_trz_is_read_by_the_reference_python_reader(tmp_path):
    """tools/py_nfclab (the reference's own TRZ reader) parses our container and returns the same frames"""
    from nfc_laboratory_b200 import export as X
    sys.path.insert(0, REF_TOOLS)
    try:
        from py_nfclab.readers import read_trz
    finally:
        sys.path.remove(REF_TOOLS)
    frames = committed_ref("test_POLL_ABF_001")[0]
    p = tmp_path / "trace.trz"
    X.write_trz(p, frames, 10_000_000)
    got = read_trz(p)
    assert len(got) == len(frames)
    for g, f in zip(got, frames):
        assert (g.tech_type, g.frame_type, g.frame_flags, g.frame_phase, g.frame_rate, g.sample_start, g.sample_end, bytes(g.frame_data)) == f
        assert g.sample_rate == 10_000_000 and g.time_start == f[5] / 10_000_000


def test_nfc_rx_lines():
    from nfc_laboratory_b200 import export as X
    reqa = (0x101, 0x102, 0x01, 0x102, 105938, 10806, 11566, b"\x52")
    o = json.loads(X.rx_json_line(reqa, 10_000_000))
    assert o == {"timestamp": 10806, "tech": "NfcA", "type": "Poll", "tech_type": 257, "frame_type": 258, "time_start": 0.0010806, "time_end": 0.0011566,
                 "sample_start": 10806, "sample_end": 11566, "sample_rate": 10000000, "date_time": 0.0010806, "rate": 105938, "data": "52", "length": 1,
                 "flags": ["request"]}
    assert X.rx_text_line(reqa, 10_000_000) == "000000.001 (Poll) [NfcA@106]: 52 "
    off = (0x100, 0x100, 0, 0x101, 0, 0, 0, b"")
    o = json.loads(X.rx_json_line(off, 10_000_000))
    assert o["time_start"] == 0 and o["date_time"] == 0 and "rate" not in o and "data" not in o and "flags" not in o and o["tech"] == "UNKNOWN"
    bad = (0x103, 0x103, 0x20 | 0x08, 0x103, 211875, 5, 9, b"\x01\x02")
    assert json.loads(X.rx_json_line(bad, 10_000_000))["flags"] == ["crc-error", "truncated", "response"]
    assert X.rx_text_line(off, 10_000_000) == "000000.000 (CarrierOff) "
