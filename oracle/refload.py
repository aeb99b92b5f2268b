"""Loads the UNMODIFIED reference (baseline/_ref, else /root/reference) next to the overlay.

Test / measurement infrastructure: only tests/, __graft_entry__.smoke() and bench.py's reference legs
import this.  The reference's modules are called `core.model.*` -- the same names as the overlay --
so they are imported with the overlay's entries temporarily removed from sys.modules / sys.path
(SURVEY 8c "purge-and-reimport"); the classes keep working afterwards because Python resolved their
cross-imports at import time.
"""
import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_PREFIXES = ("core", "cfgs", "utils")


def reference_root():
    """baseline/_ref (travels to the GPU box) or the read-only mount; None if neither exists."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("MCAN_REFERENCE", "/root/reference")):
        if cand and os.path.isdir(os.path.join(cand, "core", "model")):
            return cand
    return None


def _ours(name):
    return any(name == p or name.startswith(p + ".") for p in _PREFIXES)


@contextlib.contextmanager
def _reference_imports(ref):
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if _ours(k)}
    saved_path = list(sys.path)
    sys.path[:] = [ref] + [p for p in saved_path if os.path.abspath(p or ".") != ROOT]
    try:
        yield
    finally:
        for k in [k for k in sys.modules if _ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


class Reference(object):
    """Namespace with the reference's model modules: .net, .mca, .net_utils, .optim (+ .root)."""


_cache = {}


def load():
    """Returns the reference modules, or None when no copy of the reference is available."""
    ref = reference_root()
    if ref is None:
        return None
    if ref in _cache:
        return _cache[ref]
    with _reference_imports(ref):
        import core.model.net as net
        import core.model.mca as mca
        import core.model.net_utils as net_utils
        import core.model.optim as optim
    for m in (net, mca, net_utils, optim):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(ref)), m.__file__
    out = Reference()
    out.net, out.mca, out.net_utils, out.optim, out.root = net, mca, net_utils, optim, ref
    _cache[ref] = out
    return out


def stub_missing_modules():
    """core/exec.py and core/data/load_data.py import plotting / dataset packages that this image does
    not have: empty stand-ins are enough, the training loop never calls them."""
    def stub(name, **attrs):
        try:
            __import__(name)
            return
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)

    stub("matplotlib")
    stub("matplotlib.pyplot")
    stub("matplotlib.gridspec", GridSpec=object)
    stub("h5py")
    stub("spacy")
    stub("en_vectors_web_lg")
    stub("wandb")


def make_cfgs(ref, workdir, model_yaml, **overrides):
    """The reference's own config object (cfgs/base_cfgs.py Cfgs + cfgs/<model>.yml), built in `workdir`
    (Cfgs() creates ./results/* and ./ckpts in the current directory, cfgs/path_cfgs.py:64-77)."""
    import yaml
    cwd = os.getcwd()
    os.makedirs(os.path.join(workdir, "results"), exist_ok=True)
    os.chdir(workdir)
    try:
        with _reference_imports(ref.root):
            from cfgs.base_cfgs import Cfgs
            opt = Cfgs()
        with open(os.path.join(ref.root, "cfgs", model_yaml)) as f:
            opt.add_args(yaml.safe_load(f))
        args = {"run_mode": "train", "img_feat_pad_size": 100, "use_glove": False, "gpu": "0"}
        args.update(overrides)
        opt.add_args(args)
        saved_env = os.environ.get("CUDA_VISIBLE_DEVICES")
        opt.proc()          # derives sub_batch_size, ff_size, hidden_size_head; sets CUDA_VISIBLE_DEVICES + seeds
        if saved_env is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved_env
    finally:
        os.chdir(cwd)
    return opt
