

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


def test_wikitext_dataset_from_local_files(tmp_path):
    import pytest
    import mxb200 as mx
    from mxb200.gluon.contrib.data import text
    root = str(tmp_path)
    with pytest.raises(RuntimeError):
        text.WikiText2(root=root, segment="train")
    lines = [" = Title = ", "", " the cat sat on the mat ", " the dog sat "]
    for seg in ("train", "valid", "test"):
        with open("%s/wiki.%s.tokens" % (root, seg), "w") as f:
            f.write("\n".join(lines) + "\n")
    ds = text.WikiText2(root=root, segment="train", seq_len=4)
    vocab = ds.vocabulary
    assert "<eos>" in vocab.token_to_idx and vocab.token_to_idx["the"] > 0 and ds.frequencies["the"] == 3
    d0, l0 = ds[0]
    assert d0.shape == (4,) and l0.shape == (4,)
    toks = [t for ln in lines if ln.strip() for t in ln.split() + ["<eos>"]]
    ids = vocab.to_indices(toks)
    assert d0.asnumpy().tolist() == ids[:4] and l0.asnumpy().tolist() == ids[1:5]
    assert len(ds) == (len(ids) - 1) // 4
    val = text.WikiText2(root=root, segment="validation", vocab=vocab, seq_len=4)
    assert val.vocabulary is vocab and len(val) == len(ds)
    loader = mx.gluon.data.DataLoader(ds, batch_size=2)
    xb, yb = next(iter(loader))
    assert xb.shape == (2, 4) and yb.shape == (2, 4)
