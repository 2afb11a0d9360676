"""Executable statement of the PARALLEL TGA run-length packetiser the device runs (csrc/hana_tga.cuh), in plain Python,
next to the sequential algorithm it must reproduce byte for byte (TGAImage::unload_rle_data, tgaimage.cpp:206-246).
tests/test_tga_rle.py checks one against the other on adversarial pixel streams on the CPU box; the CUDA kernel is a
transcription of `parallel_packets` (same per-pixel formulas, the prefix quantities obtained by scans instead of loops).

Sequential algorithm as an automaton over pixels, e[i] = (pixel i == pixel i+1), e[n-1] = False:
  a packet starts; it is a RUN if e[start] else RAW;
  RAW keeps taking pixels whose e is False; a pixel with e True is left for the next packet (a run) — except that the
      128th pixel of a raw packet is taken without looking at its e (the reference's loop ends on the length first);
  RUN keeps taking pixels until it has taken one whose e is False (the last pixel of the stretch of equal pixels) or 128.

Parallel formulation. T-start: e[i] and not e[i-1]. Tail: not e[i] and e[i-1]. A "stretch" k = the pixels from one T-start
a_k up to the next: equal pixels a_k..t_k (t_k the tail), then pixels t_k+1..a_{k+1}-1 that all differ from their successor.
The only thing a stretch inherits from everything before it is one bit x_k: whether its first pixel was swallowed as the
128th pixel of the raw packet in front of it. With r = a + x, cnt = t - r + 1 (pixels of the run group), y = (cnt % 128 == 1)
(the tail is alone in its packet -> it becomes the first pixel of the raw group), rawlen = (a_next - 1 - t) + y:
      x_next = (rawlen % 128 == 127)
so x is a prefix composition of one-bit functions over the stretches; everything else is per-pixel arithmetic:
  run pixel, idx = i - r: last pixel of its packet iff idx % 128 == 127 or i == t; that pixel emits [idx % 128 + 128, B, G, R]
  raw pixel, rawidx = position in its raw group: emits [B, G, R], preceded by a header byte if rawidx % 128 == 0; the last
      pixel of a raw packet (rawidx % 128 == 127, or the group ends after it) fills the header in with rawidx % 128.
"""
import numpy as np


def sequential_packets(px):
    """px: (n,) uint32 pixel values (24 significant bits). Returns the RLE payload as bytes: the reference's algorithm."""
    n = len(px)
    out = bytearray()
    cur = 0
    while cur < n:
        ln, raw = 1, True
        while cur + ln < n and ln < 128:
            eq = px[cur + ln - 1] == px[cur + ln]
            if ln == 1:
                raw = not eq
            if raw and eq:
                ln -= 1
                break
            if not raw and not eq:
                break
            ln += 1
        out.append(ln - 1 if raw else ln + 127)
        for k in range(ln if raw else 1):
            v = int(px[cur + k])
            out += bytes((v & 255, (v >> 8) & 255, (v >> 16) & 255))
        cur += ln
    return bytes(out)


def parallel_packets(px, chunk=64):
    """The device algorithm: per-pixel roles from prefix quantities (here computed by loops chunk by chunk with an explicit
    carry, as the kernel's chained CTAs do), then scattered writes. Returns bytes."""
    n = len(px)
    px = np.asarray(px, np.uint32)
    e = np.zeros(n + 2, bool)  # e[n], e[n+1] = False: halo
    e[:n - 1] = px[:-1] == px[1:]

    def E(i):
        return bool(e[i]) if i >= 0 else False

    out = bytearray(4 * n + 16)
    # carry between chunks: last T-start, last tail (absolute, -1 = none), x of the current stretch, byte offset
    cA, cT, cX, cOff = -1, -1, 0, 0
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        # phase 1: last T-start / last tail at or before each pixel
        A, T = {}, {}
        a, t = cA, cT
        for i in range(c0, c1):
            if E(i) and not E(i - 1):
                a = i
            if (not E(i)) and E(i - 1):
                t = i
            A[i], T[i] = a, t
        # phase 2: x of the stretch each pixel lies in
        X = {}
        xc = cX  # x of the stretch the walk is in
        for i in range(c0, c1):
            X[i] = xc
            nxt_is_tstart = E(i + 1) and not E(i)
            if nxt_is_tstart:  # pixel i ends its stretch: x of the next one
                if A[i] < 0:
                    rawlen = i + 1
                else:
                    r = A[i] + xc
                    cnt = T[i] - r + 1
                    y = 1 if cnt % 128 == 1 else 0
                    rawlen = (i - T[i]) + y
                xc = 1 if rawlen % 128 == 127 else 0
        # phase 3: contributions, offsets, writes
        off = cOff
        for i in range(c0, c1):
            a, t, x = A[i], T[i], X[i]
            v = int(px[i])
            col = bytes((v & 255, (v >> 8) & 255, (v >> 16) & 255))
            # pixel i belongs to the equal-pixel part of its stretch iff e[i], or it is the tail of this stretch
            in_run_group = a >= 0 and (E(i) or ((not E(i)) and E(i - 1)))
            rawidx = None
            if a < 0:
                rawidx = i
            elif in_run_group:
                r = a + x
                if i < r:
                    rawidx = 127  # swallowed: the 128th pixel of the raw packet in front
                else:
                    idx = i - r
                    is_tail = not E(i)
                    if is_tail and (idx + 1) % 128 == 1:
                        rawidx = 0  # alone in its packet: first pixel of the raw group that follows
                    else:
                        if idx % 128 == 127 or is_tail:  # last pixel of a run packet writes it
                            out[off] = (idx % 128) + 128
                            out[off + 1:off + 4] = col
                            off += 4
                        continue
            else:
                r = a + x
                cnt = t - r + 1
                y = 1 if cnt % 128 == 1 else 0
                rawidx = (i - t - 1) + y
            k = rawidx % 128
            if k == 0:
                off += 1  # room for the header, filled in by the packet's last pixel
            out[off:off + 3] = col
            # last pixel of its raw packet?
            nxt_tstart = E(i + 1) and not E(i)  # pixel i+1 starts a stretch
            last = (k == 127) or i == n - 1 or (nxt_tstart and (k + 1) % 128 != 127)
            if last:
                out[off - 3 * k - 1] = k
            off += 3
        cA, cT, cX, cOff = A[c1 - 1], T[c1 - 1], xc, off
    return bytes(out[:cOff])
