"""Vision model zoo (reference: python/mxnet/gluon/model_zoo/vision/__init__.py)."""
from .resnet import *  # noqa: F401,F403
from .others import *  # noqa: F401,F403
from . import resnet as _resnet, others as _others


def get_model(name, **kwargs):
    models = {}
    for mod in (_resnet, _others):
        for n in mod.__all__:
            f = getattr(mod, n)
            if callable(f) and n[0].islower():
                models[n] = f
    name = name.lower()
    # the registry names of the reference (model_zoo/vision/__init__.py): 'inceptionv3', 'squeezenet1.0', 'mobilenet0.75',
    # 'mobilenetv2_1.0' ... next to the function names
    alias = {"inceptionv3": "inception_v3", "squeezenet1.0": "squeezenet1_0", "squeezenet1.1": "squeezenet1_1"}
    for w in ("1.0", "0.75", "0.5", "0.25"):
        alias["mobilenet" + w] = "mobilenet" + w.replace(".", "_")
        alias["mobilenetv2_" + w] = "mobilenet_v2_" + w.replace(".", "_")
    store_name = name                                   # the model store's file names use the registry spelling ('squeezenet1.1')
    name = alias.get(name, name)
    if name not in models:
        raise ValueError("Model %s is not supported. Available options are\n\t%s" % (name, "\n\t".join(sorted(models.keys()))))
    # pretrained / root / ctx are loader options of the reference's model zoo, not constructor arguments: pretrained weights
    # come from the local model store (model_store.get_model_file; nothing is downloaded)
    pretrained = kwargs.pop("pretrained", False)
    root = kwargs.pop("root", None)
    if pretrained:
        from ..model_store import get_model_file
        try:
            pfile = get_model_file(store_name, root=root)    # raises (with the expected location) when the file is not staged
        except Exception:  # noqa: BLE001 -- also accept the python-identifier spelling
            if store_name == name:
                raise
            pfile = get_model_file(name, root=root)
    import inspect
    fn = models[name]
    try:
        accepts_ctx = "ctx" in inspect.signature(fn).parameters or any(
            p.kind is p.VAR_KEYWORD for p in inspect.signature(fn).parameters.values())
    except (TypeError, ValueError):
        accepts_ctx = True
    if not accepts_ctx:
        kwargs.pop("ctx", None)
    try:
        net = fn(**kwargs)
        if pretrained:
            net.load_parameters(pfile, ctx=kwargs.get("ctx"))
        return net
    except TypeError as e:
        if "ctx" in str(e) and "ctx" in kwargs:
            kwargs.pop("ctx")
            return fn(**kwargs)
        raise
