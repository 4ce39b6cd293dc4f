"""Loader for pyCuSDR's JSON-with-comments config files with ``"configBase"`` includes.

The reference loads its configs with the external ``pyLoadModularJson`` package
(``pyCuSDR/pyCuSDR.py:11,61``; the tests strip comments with ``rjsmin``,
``test/loadConfig.py:37-39``).  Neither is vendored, so the same behaviour is provided here:
``//`` and ``/* */`` comments are removed outside string literals, the file named by
``"configBase"`` (relative to the including file) is loaded first, and the including file's
keys are merged over it recursively (a child value replaces the base value; dicts merge).
"""
import json
import os

__all__ = ["loadModularJson", "strip_json_comments", "merge_config"]


def strip_json_comments(text):
    out = []
    i, n = 0, len(text)
    in_str = False
    while i < n:
        c = text[i]
        if in_str:
            out.append(c)
            if c == "\\" and i + 1 < n:
                out.append(text[i + 1])
                i += 2
                continue
            if c == '"':
                in_str = False
            i += 1
        elif c == '"':
            in_str = True
            out.append(c)
            i += 1
        elif text.startswith("//", i):
            j = text.find("\n", i)
            i = n if j < 0 else j
        elif text.startswith("/*", i):
            j = text.find("*/", i + 2)
            i = n if j < 0 else j + 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


def merge_config(base, child):
    """Recursive dict merge; ``child`` wins."""
    merged = dict(base)
    for key, val in child.items():
        if isinstance(val, dict) and isinstance(merged.get(key), dict):
            merged[key] = merge_config(merged[key], val)
        else:
            merged[key] = val
    return merged


def loadModularJson(path, _depth=0):
    if _depth > 16:
        raise RecursionError("configBase include chain too deep (cycle?)")
    with open(path, "r") as f:
        conf = json.loads(strip_json_comments(f.read()))
    base_name = conf.pop("configBase", None)
    if base_name:
        base_path = os.path.join(os.path.dirname(os.path.abspath(path)), base_name)
        conf = merge_config(loadModularJson(base_path, _depth + 1), conf)
    return conf
