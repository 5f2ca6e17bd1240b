"""`Config.fromfile` for python config files with `_base_` inheritance, attribute access and `merge_from_dict`
(what tools/test.py:66-70 and configs/mcm/*.py need), plus the `DictAction` argparse action (tools/test.py:32-37)."""
import argparse
import ast
import copy
import os
import types

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"


class ConfigDict(dict):
    """dict with attribute access; nested dicts are converted on the way in."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        if isinstance(v, list):
            return [ConfigDict._wrap(x) for x in v]
        if isinstance(v, tuple):
            return tuple(ConfigDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'") from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def __deepcopy__(self, memo):
        out = ConfigDict()
        memo[id(self)] = out
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def to_dict(self):
        def un(v):
            if isinstance(v, dict):
                return {k: un(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(un(x) for x in v)
            return v
        return un(self)


def _merge(child, base):
    """child overrides base, dicts merge recursively, `_delete_=True` in a child dict replaces the base dict."""
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            if v.get(DELETE_KEY, False):
                vv = dict(v)
                vv.pop(DELETE_KEY)
                out[k] = vv
            else:
                out[k] = _merge(v, out[k])
        else:
            out[k] = copy.deepcopy(v)
    if isinstance(out, dict):
        out.pop(DELETE_KEY, None)
    return out


def _file2dict(filename):
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f'file "{filename}" does not exist')
    if not filename.endswith(".py"):
        raise OSError("this Config stand-in reads python config files only")
    with open(filename, encoding="utf-8") as f:
        text = f.read()
    ast.parse(text, filename)                  # SyntaxError with the config's own name
    scope = {"__file__": filename}
    exec(compile(text, filename, "exec"), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    if BASE_KEY in cfg:
        bases = cfg.pop(BASE_KEY)
        bases = [bases] if isinstance(bases, str) else list(bases)
        merged = {}
        for b in bases:
            bd, _ = _file2dict(os.path.join(os.path.dirname(filename), b))
            dup = merged.keys() & bd.keys()
            if dup:
                raise KeyError(f"duplicate key in base files: {sorted(dup)}")
            merged.update(bd)
        cfg = _merge(cfg, merged)
    return cfg, text


class Config:
    def __init__(self, cfg_dict=None, cfg_text=None, filename=None):
        cfg_dict = {} if cfg_dict is None else cfg_dict
        if not isinstance(cfg_dict, dict):
            raise TypeError(f"cfg_dict must be a dict, got {type(cfg_dict)}")
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)
        object.__setattr__(self, "_text", cfg_text or "")

    @staticmethod
    def fromfile(filename, use_predefined_variables=True, import_custom_modules=True):
        d, text = _file2dict(str(filename))
        return Config(d, cfg_text=text, filename=str(filename))

    @property
    def filename(self):
        return self._filename

    @property
    def text(self):
        return self._text

    @property
    def pretty_text(self):
        import pprint
        return pprint.pformat(self._cfg_dict.to_dict())

    def merge_from_dict(self, options, allow_list_keys=True):
        """options: {'a.b.c': v}: dotted keys address nested dicts (tools/test.py:68-69)."""
        nested = {}
        for full, v in options.items():
            d = nested
            keys = full.split(".")
            for k in keys[:-1]:
                d = d.setdefault(k, {})
            d[keys[-1]] = v
        object.__setattr__(self, "_cfg_dict", ConfigDict(_merge(nested, self._cfg_dict.to_dict())))

    def dump(self, file=None):
        text = self.pretty_text
        if file is None:
            return text
        with open(file, "w") as f:
            f.write(text)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    def __setitem__(self, name, value):
        self._cfg_dict[name] = value

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def __repr__(self):
        return f"Config (path: {self._filename}): {self._cfg_dict!r}"


class DictAction(argparse.Action):
    """argparse action: KEY=VALUE pairs -> dict; values parsed as int / float / bool / None / comma or bracket lists."""

    @staticmethod
    def _parse_scalar(val):
        for cast in (int, float):
            try:
                return cast(val)
            except ValueError:
                pass
        if val.lower() in ("true", "false"):
            return val.lower() == "true"
        if val == "None":
            return None
        return val

    @classmethod
    def _parse_value(cls, val):
        val = val.strip()
        if len(val) >= 2 and ((val[0] == "[" and val[-1] == "]") or (val[0] == "(" and val[-1] == ")")):
            is_tuple = val[0] == "("
            inner, items, depth, cur = val[1:-1], [], 0, ""
            for ch in inner:
                if ch in "[(":
                    depth += 1
                elif ch in "])":
                    depth -= 1
                if ch == "," and depth == 0:
                    items.append(cur)
                    cur = ""
                else:
                    cur += ch
            if cur.strip():
                items.append(cur)
            out = [cls._parse_value(i) for i in items]
            return tuple(out) if is_tuple else out
        if "," in val:
            return [cls._parse_value(v) for v in val.split(",")]
        return cls._parse_scalar(val.strip("'\""))

    def __call__(self, parser, namespace, values, option_string=None):
        options = {}
        for kv in values:
            key, val = kv.split("=", maxsplit=1)
            options[key] = self._parse_value(val)
        setattr(namespace, self.dest, options)
