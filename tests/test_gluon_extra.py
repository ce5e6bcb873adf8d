

def test_model_store_local_lookup(tmp_path):
    """pretrained=True reads from the local model store (no download): name-<sha1[:8]>.params, checksum verified."""
    import os
    import pytest
    import mxb200 as mx
    from mxb200.gluon.model_zoo import model_store, vision
    root = str(tmp_path)
    with pytest.raises(mx.MXNetError):
        model_store.get_model_file("resnet18_v1", root=root)
    net = vision.resnet18_v1(classes=10)
    net.initialize()
    net(mx.nd.zeros((1, 3, 32, 32)))
    tmp = os.path.join(root, "tmp.params")
    net.save_parameters(tmp)
    final = os.path.join(root, "resnet18_v1-%s.params" % model_store.short_hash(tmp))
    os.rename(tmp, final)
    assert model_store.get_model_file("resnet18_v1", root=root) == final
    net2 = vision.get_model("resnet18_v1", pretrained=True, root=root, classes=10)
    x = mx.nd.array(__import__("numpy").random.RandomState(0).rand(2, 3, 32, 32))
    assert abs(net(x).asnumpy() - net2(x).asnumpy()).max() < 1e-5
    # a file whose name promises another checksum is ignored
    os.rename(final, os.path.join(root, "resnet18_v1-00000000.params"))
    with pytest.raises(mx.MXNetError):
        model_store.get_model_file("resnet18_v1", root=root)
    model_store.purge(root)
    assert not [f for f in os.listdir(root) if f.endswith(".params")]
