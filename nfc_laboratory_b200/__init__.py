"""nfc_laboratory_b200 -- B200-native NFC IQ demodulation path (drop-in for lab::NfcDecoder of josevcm/nfc-laboratory).

The product is the C-ABI shared library ``libnfcb200.so`` (hand-written sm_100a CUDA, see ``csrc/`` and
``include/nfcb200.h``).  This package is the thin Python binding used by the tests and the benchmark; it holds no
decoding logic and there is no CPU fallback: importing works anywhere, decoding needs the library and a CUDA device.
"""
from .binding import (  # noqa: F401
    Frame,
    NfcB200Error,
    NfcDecoder,
    SIG_IQ_F32,
    SIG_IQ_S16,
    SIG_MAG_F32,
    SIG_MAG_S16,
    library_path,
    load_library,
)

__all__ = ["Frame", "NfcB200Error", "NfcDecoder", "SIG_IQ_F32", "SIG_MAG_F32", "SIG_MAG_S16", "SIG_IQ_S16", "library_path",
           "load_library"]
