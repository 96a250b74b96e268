"""CPU: the N4 output sink (C WAV writer of include/friendship_dispatch.h) — header fields, planar -> interleaved
frames, multi-block appends, channel mismatch.  No GPU, no oracle: the file is parsed back with the stdlib."""
import struct

import numpy as np
import pytest


def parse_wav(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    assert struct.unpack("<I", b[4:8])[0] == len(b) - 8
    pos, fmt, data, fact = 12, None, None, None
    while pos < len(b):
        tag, size = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
        body = b[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif tag == b"fact":
            fact = struct.unpack("<I", body)[0]
        elif tag == b"data":
            data = np.frombuffer(body, dtype="<f4")
        pos += 8 + size + (size & 1)
    return fmt, fact, data


@pytest.mark.parametrize("channels", [1, 2, 5])
def test_wav_client_writes_float32_frames(tmp_path, channels):
    from libfriendship_b200.dispatch import WavClient
    rng = np.random.RandomState(channels)
    blocks = [rng.uniform(-1, 1, (channels, n)).astype(np.float32) for n in (1, 4096, 5000, 0, 77)]
    path = tmp_path / "out.wav"
    w = WavClient(path, channels, 48000)
    idx = 0
    for blk in blocks:
        w.audio_rendered(blk, idx)
        idx += blk.shape[1]
    w.close()
    (tag, ch, sr, byte_rate, align, bits), fact, data = parse_wav(path)
    assert (tag, ch, sr, byte_rate, align, bits) == (3, channels, 48000, 48000 * channels * 4, channels * 4, 32)
    whole = np.concatenate(blocks, axis=1)
    assert fact == whole.shape[1] == w.frames
    assert np.array_equal(data.reshape(-1, channels).T.view(np.uint32), whole.view(np.uint32))


def test_wav_rejects_wrong_channel_count(tmp_path):
    from libfriendship_b200.dispatch import WavClient
    w = WavClient(tmp_path / "x.wav", 2, 44100)
    with pytest.raises(OSError):
        w.audio_rendered(np.zeros((3, 8), np.float32), 0)
    with pytest.raises(OSError):
        w.close()          # a failed write is reported at close too
    with pytest.raises(OSError):
        WavClient(tmp_path / "no_such_dir" / "x.wav", 1, 48000)
