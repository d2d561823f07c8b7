"""Checkpoint and wire formats around the hot path.

    load_lightning_checkpoint(model, ckpt)   reference main.py:94-121: torch.load -> ["state_dict"] -> strip the
                                             LightningModule's "model." prefix -> strict load_state_dict -> eval()
    save_weight_file / load_weight_file      the estimator's weights as one flat file in the order of libmtts' weight
                                             table (mtts_weight_name / mtts_weight_numel), so that a host without
                                             PyTorch can read it and call mtts_load_weight() entry by entry
    save_mel_npy                             one utterance's mel as (n_feats, T) float32 .npy (the hand-over format to
                                             a vocoder process; the reference itself only writes the wav, main.py:201)

Weight file layout (little endian):
    magic  8 bytes  b"MTTSW1\\0\\0"
    u32 n_entries, u32 reserved
    n_entries x { u32 name_len, name bytes (utf-8, padded with \\0 to a multiple of 8), u64 numel, u64 data_offset }
    fp32 data, every tensor 64-byte aligned at its data_offset (from the start of the file), reference layout
"""
from __future__ import annotations

import struct
from typing import Dict, Mapping, Union

import numpy as np
import torch

MAGIC = b"MTTSW1\0\0"


def strip_lightning_prefix(state_dict: Mapping[str, torch.Tensor], prefix: str = "model.") -> Dict[str, torch.Tensor]:
    """Keys of a LightningModule checkpoint carry the wrapper attribute's name (reference main.py:105-112)."""
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in state_dict.items()}


def load_lightning_checkpoint(model: torch.nn.Module, ckpt: Union[str, Mapping], map_location="cpu", strict: bool = True):
    """Load a reference training checkpoint into `model` (a matcha_tts_b200.MatchaTTS, or any module whose state-dict
    keys equal the reference's) exactly like reference main.py:94-121 and return the model in eval mode.

    ckpt: a path (torch.load, weights_only=False like the reference) or an already loaded mapping; a mapping with a
    "state_dict" entry is unwrapped.  A key mismatch raises RuntimeError like nn.Module.load_state_dict."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    sd = strip_lightning_prefix(sd)
    model.load_state_dict(sd, strict=strict)
    return model.eval()


def _estimator_table(decoder) -> Dict[str, torch.Tensor]:
    """fp32 tensors of a matcha_tts_b200.Decoder in libmtts' table order (names starting with '@' are host-derived)."""
    import math
    sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in decoder.state_dict().items()}
    half = decoder.in_channels // 2
    sd["@time_freqs"] = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))   # model.py:757-758
    from .model import estimator_param_spec
    names = ["@time_freqs"] + [k for k, _ in estimator_param_spec(decoder.in_channels, decoder.out_channels, decoder.channels[0],
                                                                 decoder.num_heads, decoder.attention_head_dim,
                                                                 decoder.num_mid_blocks)]
    return {n: sd[n] for n in names}


def save_weight_file(decoder, path: str) -> int:
    """Write the estimator weights of `decoder` as a flat weight file; returns the number of bytes written."""
    table = _estimator_table(decoder)
    head = bytearray(MAGIC + struct.pack("<II", len(table), 0))
    metas = []
    for name, t in table.items():
        nb = name.encode()
        nb += b"\0" * (-len(nb) % 8)
        metas.append((nb, t.numel()))
    head_len = len(head) + sum(4 + len(nb) + 16 for nb, _ in metas)
    off = (head_len + 63) // 64 * 64
    offs = []
    for _, n in metas:
        offs.append(off)
        off = (off + 4 * n + 63) // 64 * 64
    for (nb, n), o in zip(metas, offs):
        head += struct.pack("<I", len(nb)) + nb + struct.pack("<QQ", n, o)
    with open(path, "wb") as f:
        f.write(head)
        for t, o in zip(table.values(), offs):
            f.write(b"\0" * (o - f.tell()))
            f.write(t.numpy().astype("<f4", copy=False).tobytes())
        return f.tell()


def load_weight_file(path: str) -> Dict[str, torch.Tensor]:
    """{name: flat fp32 tensor} of a weight file, in file order."""
    buf = np.fromfile(path, dtype=np.uint8)
    if bytes(buf[:8]) != MAGIC:
        raise ValueError(f"{path}: not a matcha_tts_b200 weight file")
    n, _ = struct.unpack_from("<II", buf, 8)
    pos = 16
    out: Dict[str, torch.Tensor] = {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, pos)
        name = bytes(buf[pos + 4:pos + 4 + ln]).rstrip(b"\0").decode()
        numel, off = struct.unpack_from("<QQ", buf, pos + 4 + ln)
        pos += 4 + ln + 16
        out[name] = torch.from_numpy(buf[off:off + 4 * numel].view("<f4").copy())
    return out


def load_weight_file_into(decoder, path: str):
    """Fill a matcha_tts_b200.Decoder's parameters from a weight file (shapes from the decoder's own spec)."""
    table = load_weight_file(path)
    sd = decoder.state_dict()
    missing = [k for k in sd if k not in table]
    if missing:
        raise KeyError(f"weight file lacks {missing[:4]}...")
    decoder.load_state_dict({k: table[k].reshape(v.shape) for k, v in sd.items()}, strict=True)
    return decoder


def save_mel_npy(mel: torch.Tensor, path: str):
    """np.save of one utterance's mel as (n_feats, T) float32."""
    m = mel.detach().to("cpu", torch.float32)
    if m.ndim == 3:
        if m.shape[0] != 1:
            raise ValueError("save_mel_npy writes one utterance; index the batch first")
        m = m[0]
    np.save(path, m.numpy())
