"""Frame lists in the reference's own output formats (the step right after the hot path).
Host-side plumbing, no device code:

  trz_entry / write_frames_json   the `{"frames": [...]}` document of TraceStorageTask::writeFrameEntry
                                  (lab-tasks/src/main/cpp/tasks/TraceStorageTask.cpp:458-520), which is also the golden
                                  format of the regression tool (nfc-test/test-sdr/src/main/cpp/main.cpp writeFrames)
  write_trz / read_trz            the .trz container: tar + gzip with one member `frame.json` (README.md:378-448,
                                  readable by tools/py_nfclab/readers.py and by TraceStorageTask::readFrameEntry :380-455)
  rx_json_line / rx_text_line     one line per frame as `nfc-rx` prints them (nfc-app/app-rx/src/main/cpp/main.cpp:350-470)

A frame is anything with the fields of binding.Frame / include/nfcb200.h: tech_type, frame_type, frame_flags, frame_phase,
frame_rate, sample_start, sample_end, data -- plus sample_rate and stream_time passed by the caller (time_start =
double(sample_start) / double(sample_rate), date_time = stream_time + time_start, lab-radio NfcA.cpp:539-547).
"""
import io
import json
import math
import tarfile

FT_CARRIER_OFF, FT_CARRIER_ON, FT_POLL, FT_LISTEN = 0x0100, 0x0101, 0x0102, 0x0103
FLAG_ENCRYPTED, FLAG_TRUNCATED, FLAG_PARITY, FLAG_CRC, FLAG_SYNC = 0x02, 0x08, 0x10, 0x20, 0x40

FRAME_TYPE_NAMES = {FT_CARRIER_OFF: "CarrierOff", FT_CARRIER_ON: "CarrierOn", FT_POLL: "Poll", FT_LISTEN: "Listen"}
FRAME_TECH_NAMES = {0x0000: "None", 0x0101: "NfcA", 0x0102: "NfcB", 0x0103: "NfcF", 0x0104: "NfcV"}


def _fields(frame):
    """(tech, type, flags, phase, rate, start, end, payload) of a binding.Frame, a tuple key, or a FRAME_DTYPE record"""
    if hasattr(frame, "tech_type"):
        return (int(frame.tech_type), int(frame.frame_type), int(frame.frame_flags), int(frame.frame_phase), int(frame.frame_rate),
                int(frame.sample_start), int(frame.sample_end), bytes(frame.data))
    if hasattr(frame, "dtype") and frame.dtype.names:
        return (int(frame["tech_type"]), int(frame["frame_type"]), int(frame["frame_flags"]), int(frame["frame_phase"]), int(frame["frame_rate"]),
                int(frame["sample_start"]), int(frame["sample_end"]), bytes(frame["data"][: int(frame["length"])]))
    t = tuple(frame)
    if len(t) == 9:  # leading stream index
        t = t[1:]
    return (int(t[0]), int(t[1]), int(t[2]), int(t[3]), int(t[4]), int(t[5]), int(t[6]), bytes(t[7]))


def trz_entry(frame, sample_rate, stream_time=0.0, range_start=0.0):
    """one element of frame.json["frames"] (TraceStorageTask.cpp:462-498; key set and value types as nlohmann dumps them)"""
    tech, ftype, flags, phase, rate, start, end, payload = _fields(frame)
    time_start = float(start) / float(sample_rate)
    time_end = float(end) / float(sample_rate)
    offset = int(sample_rate * range_start)
    e = {
        "sampleStart": start - offset,
        "sampleEnd": end - offset,
        "sampleRate": int(sample_rate),
        "timeStart": time_start - range_start,
        "timeEnd": time_end - range_start,
        "techType": tech,
        "frameType": ftype,
        "frameRate": rate,
        "frameFlags": flags,
        "framePhase": phase,
        "dateTime": float(stream_time) + time_start,
    }
    if payload:
        e["frameData"] = ":".join("%02X" % b for b in payload)
        e["length"] = len(payload)
    return e


def frames_document(frames, sample_rate, stream_time=0.0, range_start=0.0, range_end=math.inf, with_length=True):
    out = []
    for f in frames:
        e = trz_entry(f, sample_rate, stream_time, range_start)
        if e["timeStart"] + range_start < range_start or e["timeEnd"] + range_start > range_end:
            continue  # TraceStorageTask.cpp:466
        if not with_length:
            e.pop("length", None)  # the regression tool's writer has no length key (test-sdr main.cpp writeFrames)
        out.append(e)
    return {"frames": out}


def write_frames_json(path, frames, sample_rate, stream_time=0.0, carrier=False):
    """the regression tool's golden file: poll / listen frames only unless carrier=True (test-sdr main.cpp:171-174)"""
    keep = [f for f in frames if carrier or _fields(f)[1] in (FT_POLL, FT_LISTEN)]
    with open(path, "w") as f:
        json.dump(frames_document(keep, sample_rate, stream_time, with_length=False), f, sort_keys=True)


def write_trz(path, frames, sample_rate, stream_time=0.0):
    """.trz = tar + gzip with the single member frame.json (ustar headers: microtar, which the reference reads TRZ with,
    knows nothing else)"""
    content = json.dumps(frames_document(frames, sample_rate, stream_time), separators=(",", ":")).encode()
    with tarfile.open(path, "w:gz", format=tarfile.USTAR_FORMAT) as tar:
        info = tarfile.TarInfo("frame.json")
        info.size = len(content)
        info.mode = 0o664
        tar.addfile(info, io.BytesIO(content))


def read_trz(path):
    """-> list of (tech, type, flags, phase, rate, sample_start, sample_end, payload) like binding.Frame.key()"""
    with tarfile.open(path, "r:gz") as tar:
        doc = json.load(tar.extractfile(tar.getmember("frame.json")))
    out = []
    for e in doc["frames"]:
        data = bytes(int(x, 16) for x in e["frameData"].split(":")) if e.get("frameData") else b""
        out.append((e["techType"], e["frameType"], e["frameFlags"], e["framePhase"], e["frameRate"], e["sampleStart"], e["sampleEnd"], data))
    return out


def rx_json_line(frame, sample_rate, stream_time=0.0):
    """nfc-rx --json line of one frame (main.cpp printFrameJSON :350-437), as compact JSON text"""
    tech, ftype, flags, phase, rate, start, end, payload = _fields(frame)
    time_start = float(start) / float(sample_rate)
    time_end = float(end) / float(sample_rate)
    date_time = float(stream_time) + time_start
    o = {
        "timestamp": start,
        "tech": FRAME_TECH_NAMES.get(tech, "UNKNOWN"),
        "type": FRAME_TYPE_NAMES.get(ftype, "UNKNOWN"),
        "tech_type": tech,
        "frame_type": ftype,
        "time_start": 0 if time_start == 0.0 else time_start,
        "time_end": 0 if time_end == 0.0 else time_end,
        "sample_start": start,
        "sample_end": end,
        "sample_rate": int(sample_rate),
        "date_time": int(date_time) if date_time == math.floor(date_time) else date_time,
    }
    if rate > 0:
        o["rate"] = rate
    if payload:
        o["data"] = ":".join("%02x" % b for b in payload)
        o["length"] = len(payload)
    fl = []
    if flags & FLAG_CRC:
        fl.append("crc-error")
    if flags & FLAG_PARITY:
        fl.append("parity-error")
    if flags & FLAG_SYNC:
        fl.append("sync-error")
    if flags & FLAG_TRUNCATED:
        fl.append("truncated")
    if flags & FLAG_ENCRYPTED:
        fl.append("encrypted")
    if ftype == FT_POLL:
        fl.append("request")
    elif ftype == FT_LISTEN:
        fl.append("response")
    if fl:
        o["flags"] = fl
    return json.dumps(o, sort_keys=True, separators=(",", ":"))  # nlohmann::json objects dump with sorted keys


def rx_text_line(frame, sample_rate):
    """nfc-rx default line of one frame (main.cpp printFrame :439-465)"""
    tech, ftype, flags, phase, rate, start, end, payload = _fields(frame)
    s = "%010.3f (%s) " % (float(start) / float(sample_rate), FRAME_TYPE_NAMES.get(ftype, "UNKNOWN"))
    if ftype in (FT_POLL, FT_LISTEN):
        import numpy as np
        khz = float(np.round(np.float32(rate) / np.float32(1000.0)))  # roundf(float(rate) / 1000.0f)
        s += "[%s@%.0f]: " % (FRAME_TECH_NAMES.get(tech, "UNKNOWN"), khz)
        s += "".join("%02X " % b for b in payload)
    return s
