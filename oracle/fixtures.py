"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Deterministic weights / inputs shared by the golden
generator and the tests, so fixtures only need to store shapes, seeds and results."""
import json
import zlib
from typing import Dict

import numpy as np
import torch


import re


def canonical_key(key: str) -> str:
    """`token_embeds[.i].proj.*` are state_dict ALIASES of the one shared `patch_embed.proj.*`
    (SURVEY.md Appendix A7): both names must carry the same tensor."""
    return re.sub(r"^token_embeds\.(\d+\.)?proj\.", "patch_embed.proj.", key)


def _seed_for(key: str, base: int) -> int:
    return (zlib.crc32(canonical_key(key).encode()) + 7919 * base) % (2 ** 31 - 1)


def det_tensor(shape, seed: int, scale: float = 1.0, offset: float = 0.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32) * scale + offset


def det_state_dict(shapes: Dict[str, tuple], base_seed: int = 0) -> Dict[str, torch.Tensor]:
    """Non-trivial deterministic parameters: norm gains ~1, biases / embeddings O(0.1),
    matrices scaled ~ 1/sqrt(fan_in) so activations stay O(1) through depth."""
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        s = _seed_for(k, base_seed)
        leaf = k.split(".")[-1]
        is_norm = ("norm" in k.split(".")[-2] if "." in k else False) or k.startswith("token_embeds.0.") or \
            k.startswith("token_embeds.2.")
        if leaf == "weight" and len(shp) == 1:
            sd[k] = det_tensor(shp, s, 0.1, 1.0)
        elif leaf == "bias" or len(shp) == 1:
            sd[k] = det_tensor(shp, s, 0.1)
        elif leaf == "weight":
            fan_in = int(np.prod(shp[1:]))
            sd[k] = det_tensor(shp, s, 1.0 / np.sqrt(fan_in))
        else:   # cls_token, pos_embed, var_embed, var_query, mask_token, decoder_pos_embed
            sd[k] = det_tensor(shp, s, 0.2)
        del is_norm
    return sd


def sd_checksum(sd) -> float:
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def save_case(path, cfg: dict, shapes: dict, arrays: dict):
    meta = {"cfg": cfg, "shapes": {k: list(v) for k, v in shapes.items()}}
    np.savez_compressed(path, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                        **{k: np.asarray(v) for k, v in arrays.items()})


def load_case(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(bytes(z["__meta__"]).decode())
    arrays = {k: z[k] for k in z.files if k != "__meta__"}
    return meta["cfg"], {k: tuple(v) for k, v in meta["shapes"].items()}, arrays
