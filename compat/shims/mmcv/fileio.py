"""mmcv.dump / mmcv.load by file extension (pickle, json, yaml): mogen/apis/test.py:108,118, tools/test.py:124."""
import json
import pickle


def _ext(path, file_format):
    if file_format is not None:
        return file_format
    return str(path).rsplit(".", 1)[-1].lower()


def dump(obj, file=None, file_format=None, **kwargs):
    fmt = _ext(file, file_format) if file is not None else file_format
    if fmt in ("pkl", "pickle"):
        if file is None:
            return pickle.dumps(obj, **kwargs)
        with open(file, "wb") as f:
            pickle.dump(obj, f, **kwargs)
    elif fmt == "json":
        if file is None:
            return json.dumps(obj, **kwargs)
        with open(file, "w") as f:
            json.dump(obj, f, **kwargs)
    elif fmt in ("yaml", "yml"):
        import yaml
        if file is None:
            return yaml.dump(obj, **kwargs)
        with open(file, "w") as f:
            yaml.dump(obj, f, **kwargs)
    else:
        raise TypeError(f"unsupported format: {fmt}")


def load(file, file_format=None, **kwargs):
    fmt = _ext(file, file_format)
    if fmt in ("pkl", "pickle"):
        with open(file, "rb") as f:
            return pickle.load(f, **kwargs)
    if fmt == "json":
        with open(file) as f:
            return json.load(f, **kwargs)
    if fmt in ("yaml", "yml"):
        import yaml
        with open(file) as f:
            return yaml.safe_load(f)
    raise TypeError(f"unsupported format: {fmt}")
