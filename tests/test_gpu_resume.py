"""GPU tests of the resumable decoder (bdf_inflate_resume_batch_{host,device}, csrc/inflate_resume.cuh) and
of the incremental DeflateDecoder mirror on it (reference src/stream.rs:243-376,
src/decompress/mod.rs:204-372).

Parity is two-fold: every stream must come out as its input whatever the cuts (zlib, the oracle's
compressor and the reference's stream tests are the producers), and every single step on the GPU
must return exactly what the host build of the same decoder core returns — status, bytes consumed,
window position, state — for a batch of decoders that stand at different places in different
streams."""
import ctypes as C
import io
import os
import random
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus  # noqa: E402
import oracle_lib as o  # noqa: E402
from host_harness import resume_host  # noqa: E402
from test_resume_host import Dribble, multi_block, plains, raw  # noqa: E402

pytestmark = pytest.mark.gpu

OK, BAD_DATA, INSUFFICIENT_SPACE, SHORT_INPUT = 0, 1, 3, 4


@pytest.fixture(scope="module")
def engine():
    import libdeflate_rsx_b200 as bdf
    return bdf


@pytest.fixture(scope="module")
def stream_mod(engine):
    from libdeflate_rsx_b200 import stream
    return stream


def read_all(dec, reads=(1 << 20,)):
    out, k = bytearray(), 0
    while True:
        chunk = dec.read(reads[k % len(reads)])
        k += 1
        if not chunk:
            return bytes(out)
        out += chunk


def test_decoder_round_trips(stream_mod):
    for p in plains():
        for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)):
            dec = stream_mod.DeflateDecoder(io.BytesIO(raw(p, level, strategy)))
            assert read_all(dec) == p
            assert dec.done


def test_decoder_cuts(stream_mod):
    rnd = random.Random(12)
    p = plains()[-1] + corpus.text_stream(7, 150000)
    comp = multi_block([p[i:i + 37000] for i in range(0, len(p), 37000)])
    for pieces, reads in (((1 << 20,), (10,)), ((571, 3, 4096), (258, 70000, 1)), ((64,), (32768,))):
        dec = stream_mod.DeflateDecoder(Dribble(comp, pieces))
        assert read_all(dec, reads) == p, (pieces, reads)
    # the oracle's compressor as the producer (levels 1, 6, 12)
    for level in (1, 6, 12):
        q = corpus.text_stream(level, 65536)
        assert read_all(stream_mod.DeflateDecoder(Dribble(o.compress(q, level, 0), (rnd.choice((100, 999)),)))) == q


def test_decoder_errors(stream_mod):
    comp = raw(corpus.text_stream(4, 100000))
    for cut in (1, 570, len(comp) // 2, len(comp) - 1):
        with pytest.raises(stream_mod.UnexpectedEof):
            stream_mod.DeflateDecoder(io.BytesIO(comp[:cut])).read_to_end()
    with pytest.raises(OSError):
        stream_mod.DeflateDecoder(io.BytesIO(b"\x07not deflate")).read()
    with pytest.raises(OSError):
        stream_mod.DeflateDecoder(io.BytesIO(comp[:300] + bytes(300) + comp[600:])).read_to_end()
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    part = c.compress(b"first part " * 100) + c.flush(zlib.Z_FULL_FLUSH)
    assert read_all(stream_mod.DeflateDecoder(io.BytesIO(part))) == b"first part " * 100       # stream.rs:367-369


def test_batch_steps_equal_the_host_core(stream_mod):
    """40 decoders over different streams, fed with different cuts, advanced together one launch per
    round: every step must equal the host build of the core bit for bit."""
    rnd = random.Random(77)
    ps = plains()
    comps = []
    for k in range(40):
        p = ps[k % len(ps)][:90000]
        comps.append(raw(p, (1, 6, 9, 0)[k % 4], zlib.Z_FIXED if k % 7 == 3 else zlib.Z_DEFAULT_STRATEGY))
    comps[5] = comps[5][:len(comps[5]) // 2]                  # truncated
    comps[6] = comps[6][:100] + bytes(50) + comps[6][150:]     # damaged
    n = len(comps)
    g_state = [bytearray(stream_mod.STATE_BYTES) for _ in range(n)]
    h_state = [bytearray(stream_mod.STATE_BYTES) for _ in range(n)]
    caps = [32768 + 258 + rnd.choice((300, 2000, 40000)) for _ in range(n)]
    g_win = [np.zeros(c, dtype=np.uint8) for c in caps]
    h_win = [np.zeros(c, dtype=np.uint8) for c in caps]
    pos = [0] * n
    fed = [0] * n
    pending = [b""] * n
    final = [False] * n
    over = [False] * n
    outs = [bytearray() for _ in range(n)]
    for rounds in range(8000):
        live = [i for i in range(n) if not over[i]]
        if not live:
            break
        for i in live:
            if caps[i] - pos[i] < 258:
                keep = 32768
                outs[i] += h_win[i][:pos[i] - keep].tobytes()
                for w in (g_win[i], h_win[i]):
                    w[:keep] = w[pos[i] - keep:pos[i]].copy()
                pos[i] = keep
        want = [resume_host.step(h_state[i], pending[i], final[i], h_win[i], pos[i]) for i in live]
        got = stream_mod.resume_step_batch([g_state[i] for i in live], [pending[i] for i in live], [final[i] for i in live],
                                           [g_win[i] for i in live], [pos[i] for i in live])
        assert got == want, rounds
        for i, (status, used, new_pos) in zip(live, want):
            assert g_state[i] == h_state[i]
            assert np.array_equal(g_win[i][:new_pos], h_win[i][:new_pos])
            pending[i] = pending[i][used:]
            pos[i] = new_pos
            if status in (OK, BAD_DATA) or (status == SHORT_INPUT and final[i]):
                over[i] = True
                outs[i] += h_win[i][:pos[i]].tobytes()
            elif status == SHORT_INPUT:
                more = comps[i][fed[i]:fed[i] + rnd.choice((7, 600, 5000, 40000))]
                fed[i] += len(more)
                pending[i] += more
                final[i] = fed[i] >= len(comps[i])
    assert all(over)
    for k in range(n):
        if k not in (5, 6):
            assert zlib.decompress(comps[k], -15) == bytes(outs[k]) == ps[k % len(ps)][:90000]


def test_device_entry(engine, stream_mod):
    import torch
    dev = torch.device("cuda", 0)
    ctx = engine.default_context()
    p = corpus.text_stream(3, 50000)
    comp = raw(p)
    n = 3
    states = torch.zeros(n * stream_mod.STATE_BYTES, dtype=torch.uint8, device=dev)
    d_in = torch.from_numpy(np.frombuffer(comp * n, dtype=np.uint8).copy()).to(dev)
    in_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * len(comp)
    fin = torch.ones(n, dtype=torch.uint8, device=dev)
    cap = 70000
    win = torch.zeros(n * cap, dtype=torch.uint8, device=dev)
    win_off = torch.arange(n, dtype=torch.int64, device=dev) * cap
    win_cap = torch.full((n,), cap, dtype=torch.int64, device=dev)
    win_pos = torch.zeros(n, dtype=torch.int64, device=dev)
    used = torch.zeros(n, dtype=torch.int64, device=dev)
    status = torch.full((n,), -1, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream(dev)
    ctx.check(ctx._lib.bdf_inflate_resume_batch_device(
        ctx.handle, n, states.data_ptr(), d_in.data_ptr(), in_off.data_ptr(), fin.data_ptr(), win.data_ptr(),
        win_off.data_ptr(), win_cap.data_ptr(), win_pos.data_ptr(), used.data_ptr(), status.data_ptr(), C.c_void_p(s.cuda_stream)))
    torch.cuda.synchronize(dev)
    assert status.tolist() == [OK] * n and win_pos.tolist() == [len(p)] * n
    for k in range(n):
        assert win[k * cap:k * cap + len(p)].cpu().numpy().tobytes() == p


def test_argument_checks(engine, stream_mod):
    st = [bytearray(stream_mod.STATE_BYTES)]
    win = [np.zeros(70000, dtype=np.uint8)]
    with pytest.raises(engine.BdfError):
        stream_mod.resume_step_batch(st, [b"abc"], [True], win, [70001])       # write position beyond the window
