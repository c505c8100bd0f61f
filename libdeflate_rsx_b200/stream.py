"""Host mirror of the reference's stream adapters (src/stream.rs) on the batch engine.

`DeflateEncoder` is the reference's encoder step for step: bytes are buffered
(1 MiB by default), a full buffer is cut into 256 KiB chunks, every chunk goes
through a fresh compressor — here: ONE `bdf_compress_units_host` call for all
chunks of the buffer — and ends in a sync flush, except the last chunk of
`finish()`, which ends the DEFLATE stream (stream.rs:42-196).  The output is
therefore byte-identical to the reference encoder's for the same sequence of
writes / flushes.

`DeflateDecoder` is the incremental reader of stream.rs:243-376: a 64 KiB window,
input pulled from the inner reader on demand, the decoder resumed where the
last read left it (`bdf_inflate_resume_batch_host`: Decompressor::decompress_streaming
for many decoder states per launch).
"""
import numpy as np

from . import _native as N
from .batch import _ptr, default_context

CHUNK = 256 * 1024


def compress_units(chunks, flush, level, context=None):
    """Compressor::compress(chunk, out, mode) for every chunk; returns a list of bytes (None = failed)."""
    ctx = context or default_context()
    n = len(chunks)
    if n == 0:
        return []
    lens = np.array([len(c) for c in chunks], dtype=np.uint64)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(bytes(c) for c in chunks) or b"\0", dtype=np.uint8)
    fl = np.array(flush, dtype=np.uint8)
    caps = lens + (lens // np.uint64(65535) + np.uint64(1)) * np.uint64(5) + np.uint64(10) + np.uint64(5)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)[:-1]
    out = np.empty(int(caps.sum()), dtype=np.uint8)
    out_size = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    ctx.check(ctx._lib.bdf_compress_units_host(ctx.handle, int(level), _ptr(flat), _ptr(off), _ptr(fl), n,
                                               _ptr(out), _ptr(out_off), _ptr(out_size), _ptr(status)))
    return [out[int(out_off[i]):int(out_off[i]) + int(out_size[i])].tobytes() if status[i] == N.OK else None
            for i in range(n)]


class DeflateEncoder:
    """io.RawIOBase-like writer: write() / flush() / finish(); also a context manager."""

    def __init__(self, writer, level, buffer_size=1024 * 1024, context=None):
        self.writer = writer
        self.level = int(level)
        self.buffer_size = int(buffer_size)
        self.buffer = bytearray()
        self.ctx = context or default_context()

    def with_buffer_size(self, size):
        self.buffer_size = int(size)
        return self

    def _flush_buffer(self, final_block):
        # flush_buffer, stream.rs:42-196
        if not self.buffer and not final_block:
            return
        data = bytes(self.buffer)
        chunks = [data[i:i + CHUNK] for i in range(0, len(data), CHUNK)] or [b""]
        flush = [N.FLUSH_SYNC] * len(chunks)
        if final_block:
            flush[-1] = N.FLUSH_FINISH
        outs = compress_units(chunks, flush, self.level, self.ctx)
        for o in outs:
            if o is None:
                raise OSError("Compression failed")
            if self.writer is not None:
                self.writer.write(o)
        self.buffer.clear()

    def write(self, buf):
        self.buffer += buf
        if len(self.buffer) >= self.buffer_size:
            self._flush_buffer(False)
        return len(buf)

    def write_all(self, buf):
        self.write(buf)

    def flush(self):
        self._flush_buffer(False)
        if self.writer is not None and hasattr(self.writer, "flush"):
            self.writer.flush()

    def finish(self):
        self._flush_buffer(True)
        w, self.writer = self.writer, None
        return w

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.writer is not None:      # Drop, stream.rs:234-240: errors are ignored there
            try:
                self._flush_buffer(True)
            except Exception:
                pass
            self.writer = None
        return False


STATE_BYTES = 368          # sizeof(bdf_inflate_state), include/bdeflate.h
PH_START, PH_STORED, PH_HUFF, PH_DONE, PH_FAILED = range(5)


def resume_step_batch(states, inputs, finals, windows, write_pos, context=None):
    """bdf_inflate_resume_batch_host for n decoders: Decompressor::decompress_streaming
    (src/decompress/mod.rs:204-372) as a batch.  states: list of writable 368-byte buffers (updated in
    place); inputs: list of bytes-like; finals: list of bool; windows: list of C-contiguous np.uint8
    arrays (written in place from write_pos[i] on).  -> list of (status, consumed, new write_pos)."""
    ctx = context or default_context()
    n = len(states)
    if n == 0:
        return []
    st = np.frombuffer(b"".join(bytes(x) for x in states), dtype=np.uint8).copy()
    lens = np.array([len(x) for x in inputs], dtype=np.uint64)
    in_off = np.zeros(n + 1, dtype=np.uint64)
    in_off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(bytes(x) for x in inputs) or b"\0", dtype=np.uint8)
    fin = np.array([1 if f else 0 for f in finals], dtype=np.uint8)
    caps = np.array([w.size for w in windows], dtype=np.uint64)
    win_off = np.zeros(n, dtype=np.uint64)
    win_off[1:] = np.cumsum(caps)[:-1]
    win = np.concatenate(windows) if n > 1 else windows[0]
    pos = np.array(write_pos, dtype=np.uint64)
    used = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    ctx.check(ctx._lib.bdf_inflate_resume_batch_host(ctx.handle, n, _ptr(st), _ptr(flat), _ptr(in_off), _ptr(fin),
                                                     _ptr(win), _ptr(win_off), _ptr(caps), _ptr(pos), _ptr(used),
                                                     _ptr(status)))
    out = []
    for i in range(n):
        states[i][:] = st[i * STATE_BYTES:(i + 1) * STATE_BYTES].tobytes()
        if n > 1:
            a, b = int(write_pos[i]), int(pos[i])
            windows[i][a:b] = win[int(win_off[i]) + a:int(win_off[i]) + b]
        out.append((int(status[i]), int(used[i]), int(pos[i])))
    return out


class UnexpectedEof(OSError, EOFError):
    """io::ErrorKind::UnexpectedEof"""


class DeflateDecoder:
    """DeflateDecoder<R: Read> (src/stream.rs:243-376): an incremental reader over a raw DEFLATE stream.

    Like the reference it keeps a 64 KiB window whose lower half is match history, pulls input from
    the inner reader only when the decoder asks for it, and resumes the decoder where the last read
    left it — here through bdf_inflate_resume_batch_host with one decoder state (the engine advances
    many states per launch; a pool of decoders would batch their steps, see resume_step_batch).
    Differences from the reference's loop follow from the engine's stop rules (include/bdeflate.h): a
    step stops when fewer than 258 bytes of room are left (not exactly 0), so the window is 64 KiB +
    258 bytes and shifts at that point; and the decoder asks for up to 570 bytes in front of a
    dynamic header, so end of input is passed down as `in_final`.
    Errors: OSError("deflate decompression failed") for invalid data (io::ErrorKind::InvalidData,
    stream.rs:330-340), UnexpectedEof (an OSError and an EOFError) for a truncated stream (:353-366).
    """

    HISTORY = 32 * 1024
    WINDOW = 64 * 1024 + 258
    INPUT_CHUNK = 32 * 1024

    def __init__(self, inner, context=None, _step=None):
        self.inner = inner
        self.ctx = context
        self._step = _step                   # tests: the host build of the same decoder core
        self.state = bytearray(STATE_BYTES)  # zeroed = start of a stream
        self.window = np.zeros(self.WINDOW, dtype=np.uint8)
        self.read_pos = 0
        self.write_pos = 0
        self.input = bytearray()
        self.in_final = False
        self.done = False
        self.steps = 0

    def _phase(self):
        return int.from_bytes(self.state[0:4], "little")

    def _one_step(self):
        self.steps += 1
        if self._step is not None:
            return self._step(self.state, bytes(self.input), self.in_final, self.window, self.write_pos)
        return resume_step_batch([self.state], [bytes(self.input)], [self.in_final], [self.window],
                                 [self.write_pos], self.ctx or default_context())[0]

    def _fill(self):
        """Decodes until the window holds unread bytes or the stream is over (the loop of stream.rs:283-375)."""
        while self.read_pos == self.write_pos and not self.done:
            if self.WINDOW - self.write_pos < 258 and self.write_pos > self.HISTORY:
                # everything has been read: keep the last 32 KiB as history (stream.rs:284-295)
                shift = self.write_pos - self.HISTORY
                self.window[:self.HISTORY] = self.window[shift:self.write_pos].copy()
                self.write_pos = self.HISTORY
                self.read_pos = self.HISTORY
            status, used, new_pos = self._one_step()
            del self.input[:used]
            self.write_pos = new_pos
            if status == N.OK:
                self.done = True
                break
            if status == N.BAD_DATA:
                raise OSError("deflate decompression failed")
            if self.write_pos > self.read_pos or status == N.INSUFFICIENT_SPACE:
                continue
            # BDF_SHORT_INPUT without progress: more input, or the end of it
            if self.in_final:
                raise UnexpectedEof("unexpected EOF")
            chunk = self.inner.read(self.INPUT_CHUNK)
            if chunk:
                self.input += chunk
            elif not self.input and self._phase() == PH_START:
                self.done = True             # the input ends between two blocks: Ok(0), stream.rs:367-369
            else:
                self.in_final = True

    def read(self, n=-1):
        """Up to n bytes (at least one unless the stream is over); n < 0: everything that is left."""
        if n is None or n < 0:
            return self.read_to_end()
        if n == 0:
            return b""
        self._fill()
        count = min(n, self.write_pos - self.read_pos)
        chunk = self.window[self.read_pos:self.read_pos + count].tobytes()
        self.read_pos += count
        return chunk

    def readinto(self, buf):
        chunk = self.read(len(buf))
        buf[:len(chunk)] = chunk
        return len(chunk)

    def read_to_end(self):
        parts = []
        while True:
            chunk = self.read(self.WINDOW)
            if not chunk:
                return b"".join(parts)
            parts.append(chunk)
