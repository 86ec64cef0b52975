"""Test infrastructure: copies `model.net.init_args`, `data` sizes and the optimizer / scheduler numbers out of every
/root/reference/configs/*/*/base_config.yaml into tests/golden/reference_yaml_configs.json, so the drop-in test
(tests/test_dropin.py) can build every shipped configuration where /root/reference does not exist (the GPU box).

    python oracle/gen_ref_configs.py
"""
import glob
import json
import os

import yaml

REF = "/root/reference/configs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_yaml_configs.json")


def load_all(root=REF):
    out = {}
    for path in sorted(glob.glob(os.path.join(root, "*", "*", "base_config.yaml"))):
        conf = yaml.safe_load(open(path))
        key = "/".join(path.split(os.sep)[-3:-1])
        data = conf.get("data", {})
        out[key] = {
            "init_args": conf["model"]["net"]["init_args"],
            "model": {k: v for k, v in conf["model"].items() if k != "net"},
            "data": {k: data[k] for k in ("batch_size", "num_classes", "single_channel", "num_channels_used", "dataset",
                                          "dict_in_variables") if k in data},
            "data_type": conf["trainer"].get("data_type"),
        }
    return out


if __name__ == "__main__":
    cfgs = load_all()
    json.dump(cfgs, open(OUT, "w"), indent=1, sort_keys=True)
    print(f"[golden] {len(cfgs)} reference YAML configs -> {OUT}")
