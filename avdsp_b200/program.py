"""Reading AVDSP program files as written by the reference's dspcreate: raw little-endian words
(`-binfile`, encoder/dsp_fileaccess.c:114-120) or a C array `const unsigned int dspFactory[] = {0x..., };`
(`-hexfile`, encoder/dspcreate.c:23-24, encoder/dsp_fileaccess.c:122-134)."""
from __future__ import annotations

import re

import numpy as np

FREQS = (8000, 16000, 24000, 32000, 44100, 48000, 88200, 96000,
         176400, 192000, 352800, 384000, 705600, 768000)     # runtime/dsp_header.h:136-145


def load_bin(path) -> np.ndarray:
    raw = open(path, "rb").read()
    return np.frombuffer(raw[: len(raw) // 4 * 4], dtype="<u4").view(np.int32).copy()


def load_hex(path) -> np.ndarray:
    text = open(path, "r").read()
    body = text[text.index("{") + 1: text.rindex("}")] if "{" in text else text
    vals = [int(t, 16) for t in re.findall(r"0[xX]([0-9a-fA-F]+)", body)]
    return np.array(vals, dtype=np.uint32).view(np.int32)


def load(path) -> np.ndarray:
    p = str(path)
    return load_hex(p) if p.endswith((".hex", ".h", ".c")) else load_bin(p)


def header(words) -> dict:
    """The 12-word dspHeader_t (runtime/dsp_header.h:213-228)."""
    w = [int(x) & 0xFFFFFFFF for x in words[:12]]
    return dict(totalLength=w[1], dataSize=w[2], checkSum=w[3], numCores=w[4], version=w[5],
                encoding=w[6] & 0xFFFF, maxOpcode=w[6] >> 16, freqMin=w[7], freqMax=w[8],
                usedInputs=w[9], usedOutputs=w[10], serialHash=w[11])
