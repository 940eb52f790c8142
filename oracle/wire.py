"""AVDSP wire format helpers for the TEST side: opcode table, disassembler and a tiny assembler.

TEST INFRASTRUCTURE ONLY (lives under oracle/).  The product has its own decoder in
avdsp_b200/csrc/decoder.cpp; nothing there imports this.

Wire format facts restated from /root/reference/module_avdsp/runtime/dsp_header.h:40-132 (opcode
enum), :197-228 (opcode word, header) and encoder/dsp_encoder.c (layouts of the PARAM payloads:
biquad sections :1225-1290, LOAD_MUX :798-813, delays :1088-1160, FIR :1292-1372).

The assembler exists because (a) the reference encoder cannot emit a valid DSP_FIR program
 and (b) fuzz tests want arbitrary opcode mixes.  It emits what the
*runtime* decodes (dsp_runtime.c:302-1314), which is the contract the executor implements.
"""
from __future__ import annotations

import math
import struct
from typing import Iterable, List, Sequence

import numpy as np

OPCODES = [
    "END_OF_CODE", "HEADER", "NOP", "CORE", "PARAM", "PARAM_NUM", "SERIAL",
    "TPDF_CALC", "TPDF", "WHITE", "CLRXY", "SWAPXY", "COPYXY", "COPYYX",
    "ADDXY", "ADDYX", "SUBXY", "SUBYX", "MULXY", "DIVXY", "DIVYX", "AVGXY", "AVGYX",
    "NEGX", "NEGY", "SQRTX", "SHIFT", "VALUE", "VALUE_INT", "MUL_VALUE", "MUL_VALUE_INT",
    "DIV_VALUE", "DIV_VALUE_INT", "AND_VALUE_INT",
    "LOAD", "LOAD_GAIN", "LOAD_MUX", "STORE", "LOAD_STORE", "LOAD_MEM", "STORE_MEM",
    "GAIN", "SAT0DB", "SAT0DB_TPDF", "SAT0DB_GAIN", "SAT0DB_TPDF_GAIN",
    "DELAY_1", "DELAY", "DELAY_DP", "DATA_TABLE", "BIQUADS", "FIR",
    "RMS", "DCBLOCK", "DITHER", "DITHER_NS2", "DISTRIB", "DIRAC", "SQUAREWAVE", "CLIP",
    "LOAD_MEM_DATA", "SINE",
]
OP = {n: i for i, n in enumerate(OPCODES)}
MAX_OPCODE = len(OPCODES)
FREQS = [8000, 16000, 24000, 32000, 44100, 48000, 88200, 96000,
         176400, 192000, 352800, 384000, 705600, 768000]
MANT = 28
ENCODER_VERSION = 0x102


def freq_index(fs: int) -> int:
    return FREQS.index(fs)


def load_bin(path) -> np.ndarray:
    raw = open(path, "rb").read()
    return np.frombuffer(raw[: len(raw) // 4 * 4], dtype="<i4").copy()


def f32_bits(x: float) -> int:
    return struct.unpack("<i", struct.pack("<f", float(x)))[0]


def q28(x: float) -> int:
    """DSP_QM32(x,28): truncating conversion with saturation (dsp_header.h:276-285)."""
    if x >= 8.0:
        return 0x7FFFFFFF
    if x < -8.0:
        return -0x80000000
    return int(x * (1 << 28))  # C cast truncates toward zero, so does int()


def header(words: Sequence[int]) -> dict:
    w = [int(x) for x in words[:12]]
    return dict(op=(w[0] >> 16) & 0xFFFF, skip=w[0] & 0xFFFF, totalLength=w[1], dataSize=w[2],
                checkSum=w[3] & 0xFFFFFFFF, numCores=w[4], version=w[5], format=w[6] & 0xFFFF,
                maxOpcode=(w[6] >> 16) & 0xFFFF, freqMin=w[7], freqMax=w[8],
                usedInputs=w[9] & 0xFFFFFFFF, usedOutputs=w[10] & 0xFFFFFFFF, serialHash=w[11] & 0xFFFFFFFF)


def walk(words: Sequence[int]) -> Iterable[tuple]:
    """Yield (index, opcode, skip) for every opcode word, like dspCalcSumCore (dsp_header.h:234-251)."""
    p = 0
    n = len(words)
    while p < n:
        w = int(words[p]) & 0xFFFFFFFF
        op, skip = w >> 16, w & 0xFFFF
        yield p, op, skip
        if skip == 0:
            return
        p += skip


def checksum(words: Sequence[int]) -> tuple:
    s, cores = 0, 0
    for p, op, skip in walk(words):
        if skip == 0:
            break
        if op == OP["CORE"]:
            cores += 1
        s = (s + (int(words[p]) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return s, max(cores, 1)


def disassemble(words: Sequence[int]) -> List[str]:
    out = []
    for p, op, skip in walk(words):
        name = OPCODES[op] if op < MAX_OPCODE else f"?{op}"
        args = [int(x) for x in words[p + 1: p + min(skip, 6)]] if op not in (OP["PARAM"], OP["PARAM_NUM"]) else []
        out.append(f"{p:5d} {name:<18s} skip={skip:<4d} {args}")
    return out


def io_maps(words: Sequence[int]):
    """Union of the DSP_CORE used-input / used-output bitmaps (encoder :454-462, :624-632)."""
    ins, outs = 0, 0
    seen = False
    for p, op, skip in walk(words):
        if skip == 0:
            break
        if op == OP["CORE"]:
            seen = True
            ins |= int(words[p + 1]) & 0xFFFFFFFF
            outs |= int(words[p + 2]) & 0xFFFFFFFF
    if not seen:
        h = header(words)
        ins, outs = h["usedInputs"], h["usedOutputs"]
    return [i for i in range(32) if ins >> i & 1], [i for i in range(32) if outs >> i & 1]


class Asm:
    """Minimal AVDSP program assembler (runtime-decodable layouts).  Values are given in the
    encoding of `fmt`: Q4.28 ints for fmt 2, IEEE float bits otherwise."""

    def __init__(self, fmt: int = 2, fmin: int = 48000, fmax: int = 48000):
        self.fmt = fmt
        self.fmin, self.fmax = freq_index(fmin), freq_index(fmax)
        self.nf = self.fmax - self.fmin + 1
        self.w: List[int] = [0] * 12
        self.data = 0
        self.open_op = None       # index of an opcode whose skip is not yet known
        self.core_at = None
        self.ins = self.outs = 0
        self.cins = self.couts = 0
        self.maxop = 0

    # -- low level ---------------------------------------------------------------------------
    def here(self) -> int:
        return len(self.w)

    def word(self, v: int) -> int:
        self.w.append(int(v) & 0xFFFFFFFF)
        return len(self.w) - 1

    def num(self, x: float) -> int:
        return self.word(q28(x) if self.fmt == 2 else f32_bits(x))

    def _close(self):
        if self.open_op is not None:
            self.w[self.open_op] = (self.w[self.open_op] & 0xFFFF0000) | ((len(self.w) - self.open_op) & 0xFFFF)
            self.open_op = None

    def op(self, name: str) -> int:
        self._close()
        code = OP[name]
        self.maxop = max(self.maxop, code)
        self.open_op = self.word(code << 16)
        return self.open_op

    def alloc(self, n: int, align8: bool = False, misalign8: bool = False) -> int:
        if align8 and (self.data & 1):
            self.data += 1
        if misalign8 and not (self.data & 1):
            self.data += 1
        a = self.data
        self.data += n
        return a

    # -- structure ---------------------------------------------------------------------------
    def core(self):
        self._flush_core()
        at = self.op("CORE")
        self.word(0)
        self.word(0)
        self.core_at = at
        self.cins = self.couts = 0

    def _flush_core(self):
        if self.core_at is not None:
            self.w[self.core_at + 1] = self.cins
            self.w[self.core_at + 2] = self.couts
            self.core_at = None

    def param(self) -> int:
        return self.op("PARAM")

    def nop(self):
        self.op("NOP")

    def align_param(self, odd: bool):
        """pad inside a PARAM so that the next word lands on an odd/even index"""
        if (self.here() & 1) != (1 if odd else 0):
            self.word(0)

    # -- PARAM payloads ------------------------------------------------------------------------
    def biquad_sections(self, coefs_per_fs: Sequence[Sequence[Sequence[float]]], bypass_flag: int = 1) -> int:
        """coefs_per_fs[section][fs_index] = (b0, b1, b2, a1 (not yet reduced), a2).  Returns header index."""
        self.align_param(odd=True)
        h = self.word((OP["BIQUADS"] << 16) | len(coefs_per_fs))
        self.word(bypass_flag)
        for sec in coefs_per_fs:
            assert len(sec) == self.nf
            self.word(0)              # type<<16 | freq (informational)
            self.word(f32_bits(0.0))  # Q
            self.word(f32_bits(1.0))  # gain
            for k, (b0, b1, b2, a1, a2) in enumerate(sec):
                if self.here() & 1:
                    self.word(0)
                for c in (b0, b1, b2, a1 - 1.0, a2):
                    self.num(c)
        return h

    def mux_table(self, pairs: Sequence[tuple]) -> int:
        h = self.word((OP["LOAD_MUX"] << 16) | len(pairs))
        for io, g in pairs:
            self.word(io)
            self.num(g)
            self._in(io)
        return h

    def delay_param(self, max_us: int, us: int, fs_max: int) -> int:
        max_samples = (max_us * fs_max + 500000) // 1000000
        return self.word((max_samples << 16) | (us & 0xFFFF))

    def mem_location(self, init=(0, 0)) -> int:
        if self.here() & 1:
            self.word(0)
        a = self.word(init[0])
        self.word(init[1])
        return a

    def fir_impulses(self, impulses: Sequence) -> List[int]:
        """One entry per covered fs: a list/array of taps, ('delay', n) or None (-> not run).
        Returns the word index of each impulse's length word."""
        assert len(impulses) == self.nf
        where = []
        for imp in impulses:
            self.align_param(odd=True)     # length word odd => taps 8-byte aligned
            if imp is None:
                where.append(0)
                continue
            if isinstance(imp, tuple) and imp[0] == "delay":
                where.append(self.word(int(imp[1]) << 16))
                self.word(0)
                continue
            taps = list(imp)
            where.append(self.word(len(taps)))
            for t in taps:
                self.num(t) if self.fmt != 2 else self.word(int(t))
        return where

    # -- opcodes -------------------------------------------------------------------------------
    def _in(self, io):
        if io < 32:
            self.ins |= 1 << io
            self.cins |= 1 << io

    def _out(self, io):
        if io < 32:
            self.outs |= 1 << io
            self.couts |= 1 << io

    def simple(self, name: str):
        self.op(name)

    def load(self, io):
        self.op("LOAD"); self.word(io); self._in(io)

    def store(self, io):
        self.op("STORE"); self.word(io); self._out(io)

    def load_gain(self, io, g):
        self.op("LOAD_GAIN"); self.word(io); self.word(3); self.num(g); self._in(io)

    def _ptr_op(self, name, addr, value=None):
        at = self.op(name)
        if addr:
            self.word(addr - at)
        else:
            self.word(2)
            self.num(value)

    def gain(self, g=None, addr=0):
        self._ptr_op("GAIN", addr, g)

    def sat0db(self):
        self.op("SAT0DB")

    def sat0db_tpdf(self):
        self.op("SAT0DB_TPDF")

    def sat0db_gain(self, g=None, addr=0):
        self._ptr_op("SAT0DB_GAIN", addr, g)

    def sat0db_tpdf_gain(self, g=None, addr=0):
        self._ptr_op("SAT0DB_TPDF_GAIN", addr, g)

    def value(self, v):
        self._ptr_op("VALUE", 0, v)

    def value_int(self, v):
        at = self.op("VALUE_INT"); self.word(2); self.word(v)

    def imm(self, name, v, as_int=False):
        self.op(name)
        self.word(v) if as_int else self.num(v)

    def shift(self, n):
        self.op("SHIFT"); self.word(n)

    def tpdf_calc(self, dither) -> int:
        self.op("TPDF_CALC"); self.word(dither)
        a = self.alloc(2, align8=True); self.word(a); return a

    def tpdf(self, dither) -> int:
        self.op("TPDF"); self.word(dither)
        a = self.alloc(2, align8=True); self.word(a); return a

    def load_mux(self, table_addr) -> int:
        at = self.op("LOAD_MUX"); self.word(table_addr - at)
        a = self.alloc(2, align8=True); self.word(a); return a

    def load_store(self, pairs):
        self.op("LOAD_STORE")
        for i, o in pairs:
            self.word(i); self.word(o); self._in(i); self._out(o)

    def load_mem(self, addr):
        at = self.op("LOAD_MEM"); self.word(addr - at)

    def store_mem(self, addr):
        at = self.op("STORE_MEM"); self.word(addr - at)

    def load_mem_data(self, data_off):
        self.op("LOAD_MEM_DATA"); self.word(data_off)

    def delay_1(self):
        self.op("DELAY_1"); self.word(self.alloc(2, align8=True))

    def delay_fixed_us(self, us, fs_max, dp=False):
        self.op("DELAY_DP" if dp else "DELAY")
        factor = int(4294.967296 * fs_max) & 0xFFFFFFFF
        max_samples = (factor * us) >> 32
        self.word(us)
        self.word(self.alloc(1 + max_samples * 2, misalign8=True) if dp else self.alloc(1 + max_samples))
        self.word(0)

    def delay(self, param_addr, dp=False):
        at = self.op("DELAY_DP" if dp else "DELAY")
        size = (self.w[param_addr] >> 16) & 0xFFFF
        self.word(size)
        self.word(self.alloc(size * 2 + 1, misalign8=True) if dp else self.alloc(size + 1))
        self.word(param_addr - at)

    def biquads(self, header_addr):
        at = self.op("BIQUADS")
        num = self.w[header_addr] & 0xFFFF
        self.word(self.alloc(num * 6, align8=True))
        self.word(header_addr - at)

    def fir(self, impulse_addrs: Sequence[int], max_len: int):
        at = self.op("FIR")
        for a in impulse_addrs:
            self.word(a - at if a else 0)
        self.word(self.alloc(max_len, align8=True))

    def dcblock(self, poles: Sequence[float]):
        self.op("DCBLOCK"); self.word(self.alloc(4, align8=True))
        for p in poles:
            self.num(p)

    def dither(self):
        self.op("DITHER"); self.word(self.alloc(6, align8=True))

    def dither_ns2(self, table_addr):
        at = self.op("DITHER_NS2"); self.word(self.alloc(3)); self.word(table_addr - at)

    def clip(self, v):
        self.op("CLIP"); self.num(v)

    def dirac(self, gain, counts: Sequence[int], square=False):
        self.op("SQUAREWAVE" if square else "DIRAC"); self.word(self.alloc(1)); self.num(gain)
        for c in counts:
            self.word(c)

    def data_table(self, gain, div, size, table_addr):
        at = self.op("DATA_TABLE"); self.num(gain); self.word(div); self.word(size)
        self.word(self.alloc(1)); self.word(table_addr - at)

    # -- finish ----------------------------------------------------------------------------------
    def end(self) -> np.ndarray:
        self._flush_core()
        self._close()
        self.w.append(0)                      # END_OF_CODE
        if len(self.w) & 1:
            self.w.append(0)
        self.w[0] = (OP["HEADER"] << 16) | 12
        self.w[1] = len(self.w)
        self.w[2] = self.data
        self.w[5] = ENCODER_VERSION
        self.w[6] = (self.maxop << 16) | (MANT if self.fmt == 2 else 0)
        self.w[7], self.w[8] = self.fmin, self.fmax
        self.w[9], self.w[10] = self.ins, self.outs
        self.w[11] = 0
        s, cores = checksum(self.w)
        self.w[3] = s
        self.w[4] = cores
        return np.array(self.w, dtype=np.uint32).view(np.int32)


def rbj_peak(fs, f, q, gain):
    """Same formulas as dspFilter2ndOrder(FPEAK) (encoder/dsp_filters.c:140-148); used by fuzz programs."""
    w0 = 2 * math.pi * f / fs
    alpha = math.sin(w0) / 2 / q
    A = math.sqrt(gain)
    a0 = 1 + alpha / A
    return ((1 + alpha * A) / a0, -2 * math.cos(w0) / a0, (1 - alpha * A) / a0,
            2 * math.cos(w0) / a0, -(1 - alpha / A) / a0)
