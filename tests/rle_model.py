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


# ----------------------------------------------------------------------------------------------------------------------
# The three-kernel form of csrc/hana_tga.cuh (round 2, second version): (A) e-bits per pixel packed 32 to a word,
# (B) ONE thread block per frame walks the words — each thread a contiguous span of them — and leaves, per word, the
# state at its first pixel (last T-start a, last tail t, x of the current stretch, byte offset) plus the x of every
# T-start inside it; (C) one warp per word assigns roles per pixel from that record with the formulas above and writes.
# `words_packets` states B and C in Python (B with the word-level shortcuts the kernel takes), word by word.
def words_packets(px, span=4):
    n = len(px)
    px = np.asarray(px, np.uint32)
    nw = (n + 31) // 32
    ebits = [0] * (nw + 1)
    for i in range(n - 1):
        if px[i] == px[i + 1]:
            ebits[i >> 5] |= 1 << (i & 31)

    def word(w):
        return ebits[w] if 0 <= w <= nw else 0

    # ---- B, pass 0: last T-start / last tail in front of every span of `span` words
    def masks(w):
        cur, eprev, enext = word(w), (word(w - 1) >> 31) & 1 if w > 0 else 0, word(w + 1) & 1
        sh = ((cur << 1) | eprev) & 0xFFFFFFFF          # bit j = e(j-1)
        up = ((cur >> 1) | (enext << 31)) & 0xFFFFFFFF  # bit j = e(j+1)
        ts = cur & ~sh & 0xFFFFFFFF                     # T-starts
        tl = ~cur & sh & 0xFFFFFFFF                     # tails
        se = ~cur & up & 0xFFFFFFFF                     # stretch ends (the next pixel is a T-start)
        return cur, eprev, ts, tl, se

    def hi(m):
        return m.bit_length() - 1

    nspans = (nw + span - 1) // span
    span_a, span_t = [-1] * nspans, [-1] * nspans
    for s in range(nspans):
        for w in range(s * span, min(nw, (s + 1) * span)):
            _, _, ts, tl, _ = masks(w)
            if ts:
                span_a[s] = w * 32 + hi(ts)
            if tl:
                span_t[s] = w * 32 + hi(tl)
    a_in, t_in = [-1] * nspans, [-1] * nspans
    for s in range(1, nspans):
        a_in[s] = max(a_in[s - 1], span_a[s - 1])
        t_in[s] = max(t_in[s - 1], span_t[s - 1])

    # ---- B, pass 1: x -> x' of every span (events of a word in pixel order)
    def walk_x(s, x, record=None):
        a, t = a_in[s], t_in[s]
        for w in range(s * span, min(nw, (s + 1) * span)):
            cur, eprev, ts, tl, se = masks(w)
            if record is not None:
                record[w] = [a, t, x, 0]
            ev = ts | tl | se
            while ev:
                j = (ev & -ev).bit_length() - 1
                ev &= ev - 1
                i = w * 32 + j
                if (ts >> j) & 1:
                    a = i
                    if record is not None and j > 0:
                        record[w][3] |= x << j  # x of a T-start inside the word (bit 0: the word's own x)
                if (tl >> j) & 1:
                    t = i
                if (se >> j) & 1:
                    rawlen = i + 1 if a < 0 else (i - t) + (1 if ((t - a - x + 1) & 127) == 1 else 0)
                    x = 1 if (rawlen & 127) == 127 else 0
        return x

    F = [(walk_x(s, 0), walk_x(s, 1)) for s in range(nspans)]
    x_in = [0] * nspans
    for s in range(1, nspans):
        x_in[s] = F[s - 1][x_in[s - 1]]
    # ---- B, pass 2: per-word records + byte counts
    rec = {}
    for s in range(nspans):
        walk_x(s, x_in[s], rec)

    def pixel_role(i, a, t, x):
        """(role, k, last): role 0 nothing, 1 last pixel of a run packet (k = idx % 128), 2 raw (k = rawidx % 128)"""
        def E(q):
            return (word(q >> 5) >> (q & 31)) & 1 if 0 <= q < n else 0
        is_e, is_tail = E(i), (not E(i)) and E(i - 1)
        nxt_tstart = E(i + 1) and not is_e
        rawidx = None
        if a < 0:
            rawidx = i
        elif is_e or is_tail:
            r = a + x
            if i < r:
                rawidx = 127
            else:
                idx = i - r
                if is_tail and ((idx + 1) & 127) == 1:
                    rawidx = 0
                elif (idx & 127) == 127 or is_tail:
                    return 1, idx & 127, True
                else:
                    return 0, 0, False
        else:
            r = a + x
            rawidx = (i - t - 1) + (1 if ((t - r + 1) & 127) == 1 else 0)
        k = rawidx & 127
        last = k == 127 or i == n - 1 or (nxt_tstart and ((k + 1) & 127) != 127)
        return 2, k, last

    def lane_state(w, j):
        """what lane j of word w's warp derives from the word record: (a, t, x) of pixel w*32+j"""
        a0, t0, x0, xmask = rec[w]
        cur, eprev, ts, tl, _ = masks(w)
        m = ts & ((2 << j) - 1)
        a = w * 32 + hi(m) if m else a0
        x = (xmask >> hi(m)) & 1 if m and hi(m) > 0 else x0  # a T-start at bit 0 begins the stretch the word starts in
        mt = tl & ((2 << j) - 1)
        t = w * 32 + hi(mt) if mt else t0
        return a, t, x

    def word_bytes(w):
        a0, t0, x0, _ = rec[w]
        cur, eprev, ts, tl, se = masks(w)
        full = (w + 1) * 32 <= n
        if full and cur == 0xFFFFFFFF and eprev:  # shortcut: the middle of a run
            r = a0 + x0
            return 4 if ((127 - (w * 32 - r)) & 127) < 32 else 0
        if full and cur == 0 and not eprev and not (word(w + 1) & 1):  # shortcut: the middle of a raw stretch
            rawidx0 = w * 32 if a0 < 0 else (w * 32 - t0 - 1) + (1 if ((t0 - (a0 + x0) + 1) & 127) == 1 else 0)
            return 96 + (1 if ((-rawidx0) & 127) < 32 else 0)
        tot = 0
        for j in range(min(32, n - w * 32)):
            role, k, _ = pixel_role(w * 32 + j, *lane_state(w, j))
            tot += 4 if role == 1 else (3 + (1 if k == 0 else 0)) if role == 2 else 0
        return tot

    offs, o = [], 0
    for w in range(nw):
        offs.append(o)
        o += word_bytes(w)
    # ---- C: one warp per word
    out = bytearray(o + 8)
    for w in range(nw):
        pos = offs[w]
        for j in range(min(32, n - w * 32)):
            i = w * 32 + j
            role, k, last = pixel_role(i, *lane_state(w, j))
            v = int(px[i])
            col = bytes((v & 255, (v >> 8) & 255, (v >> 16) & 255))
            if role == 1:
                out[pos] = k + 128
                out[pos + 1:pos + 4] = col
                pos += 4
            elif role == 2:
                cpos = pos + (1 if k == 0 else 0)
                out[cpos:cpos + 3] = col
                if last:
                    out[cpos - 3 * k - 1] = k
                pos = cpos + 3
    return bytes(out[:o])
