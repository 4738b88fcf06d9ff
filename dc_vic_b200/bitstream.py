"""DC-VIC's .bin wire format (src/utils/codec_utils.py:16-66), host side: a 6-byte header
(uint16 H, uint16 W, uint8 max|y_hat|, uint8 quality index) and uint32-length-prefixed strings
[header, z string, y string] (hyperprior_dc_vic_model.py:330-376, scripts/compress.py).  Byte shuffling of a few
hundred bytes: numpy, no device work."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

__all__ = ["HeaderHandler", "save_byte_strings", "load_byte_strings", "pack_byte_strings", "unpack_byte_strings"]


class HeaderHandler:
    """codec_utils.py:7-47."""

    @staticmethod
    def check_img_size(img_size) -> None:
        assert len(img_size) == 2 and isinstance(img_size[0], int) and isinstance(img_size[1], int)

    def encode(self, img_size: Tuple[int, int], y_hat: torch.Tensor, quality_ind: int) -> bytes:
        self.check_img_size(img_size)
        max_val = int(torch.max(torch.abs(y_hat)))
        return (np.array(list(img_size), dtype=np.uint16).tobytes() + np.array(max_val, dtype=np.uint8).tobytes()
                + np.array(quality_ind, dtype=np.uint8).tobytes())

    def decode(self, header_byte_string: bytes) -> Dict:
        img_size = np.frombuffer(header_byte_string[:4], dtype=np.uint16)
        return {"img_size": (int(img_size[0]), int(img_size[1])),
                "max_sample": int(np.frombuffer(header_byte_string[4:5], dtype=np.uint8)[0]),
                "quality_ind": int(np.frombuffer(header_byte_string[5:6], dtype=np.uint8)[0])}


def pack_byte_strings(string_list: Sequence[bytes]) -> bytes:
    return b"".join(np.array(len(s), dtype=np.uint32).tobytes() + s for s in string_list)


def unpack_byte_strings(blob: bytes) -> List[bytes]:
    out, p = [], 0
    while p < len(blob):
        n = int(np.frombuffer(blob[p:p + 4], dtype=np.uint32)[0])
        out.append(blob[p + 4:p + 4 + n])
        p += 4 + n
    return out


def save_byte_strings(save_path: str, string_list: Sequence[bytes]) -> None:
    with open(save_path, "wb") as f:
        f.write(pack_byte_strings(string_list))


def load_byte_strings(load_path: str) -> List[bytes]:
    with open(load_path, "rb") as f:
        return unpack_byte_strings(f.read())
