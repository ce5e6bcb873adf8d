#!/usr/bin/env python
"""Elastic ImageNet training (the fork's headline example, reference example/dynamic-training/train_resnet.py) on mxb200.

Same command line and the same three pieces that make the reference example elastic:
  * ``ElasticImageRecordIter`` : a ``mx.mod.BaseDataIterator`` whose ``get_data_iterator(kv)`` re-creates the record
    iterator for the CURRENT (kv.num_workers, kv.rank) at every epoch boundary -- the global batch stays constant, the
    per-worker batch = global / num_workers changes;
  * ``ElasticSpeedometer``     : throughput logging that follows the changing per-worker batch;
  * ``Module.fit(..., data_iterator=...)`` : the elastic epoch loop (membership-change barrier, newcomers start at
    EPOCH_BEGIN with parameters pulled from the servers).

Launch (reference example/dynamic-training/run.sh):
  python tools/launch.py -n $N -H hosts --elastic-training-enabled True \\
      python examples/dynamic-training/train_resnet.py --use-rec --batch-size 4096 --dtype bfloat16 --num-epochs 90 \\
      --lr 1.6 --lr-mode poly --warmup-epochs 5 --warmup-lr 0.001 --last-gamma --mode symbolic --model resnet50_v1b \\
      --kvstore dist_sync_device --rec-train train.rec --rec-train-idx train.idx --rec-val val.rec --rec-val-idx val.idx
``--benchmark 1`` replaces the record files by synthetic data (no dataset on this box).
``--mode hybrid --kvstore elastic_device`` is the B200-native variant: one process per GPU on one box, the fused P2P
allreduce+update kernel instead of parameter servers, same host-file driven membership changes.
"""
import argparse
import logging
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mxb200 as mx  # noqa: E402
from mxb200 import autograd as ag, gluon  # noqa: E402
from mxb200.gluon.model_zoo import vision  # noqa: E402

parser = argparse.ArgumentParser(description='Train a model for image classification (elastic).')
parser.add_argument('--data-dir', type=str, default='~/.mxnet/datasets/imagenet')
parser.add_argument('--rec-train', type=str, default='')
parser.add_argument('--rec-train-idx', type=str, default='')
parser.add_argument('--rec-val', type=str, default='')
parser.add_argument('--rec-val-idx', type=str, default='')
parser.add_argument('--use-rec', action='store_true')
parser.add_argument('--benchmark', type=int, default=0, help='1 = synthetic data')
parser.add_argument('--batch-size', type=int, default=128, help='GLOBAL batch size, split over the live workers')
parser.add_argument('--dtype', type=str, default='float32')
parser.add_argument('--gpus', type=str, default='0')
parser.add_argument('-j', '--num-data-workers', dest='num_workers', default=4, type=int)
parser.add_argument('--num-epochs', type=int, default=90)
parser.add_argument('--lr', type=float, default=6.4)
parser.add_argument('--momentum', type=float, default=0.9)
parser.add_argument('--wd', type=float, default=0.0001)
parser.add_argument('--lr-mode', type=str, default='step')
parser.add_argument('--lr-poly-power', type=int, default=2)
parser.add_argument('--lr-decay', type=float, default=0.1)
parser.add_argument('--lr-decay-epoch', type=str, default='40,60')
parser.add_argument('--warmup-lr', type=float, default=0.001)
parser.add_argument('--warmup-epochs', type=int, default=0)
parser.add_argument('--last-gamma', action='store_true')
parser.add_argument('--mode', type=str, default='symbolic', help='symbolic | imperative | hybrid')
parser.add_argument('--model', type=str, required=True)
parser.add_argument('--use-pretrained', action='store_true')
parser.add_argument('--log-interval', type=int, default=50)
parser.add_argument('--save-frequency', type=int, default=0)
parser.add_argument('--save-dir', type=str, default='params')
parser.add_argument('--logging-dir', type=str, default='logs')
parser.add_argument('--kvstore', type=str, default='device')
parser.add_argument('--image-size', type=int, default=224)
parser.add_argument('--num-classes', type=int, default=1000)
parser.add_argument('--num-training-samples', type=int, default=1281167)
parser.add_argument('--dynamic-minibatch', type=str, default='constant_global',
                    help='constant_global (reference behaviour: global batch fixed, no lr change) | linear | sqrt: lr rule of the '
                         'dynamic mini-batch scheduler when the PER-WORKER batch is kept and the global batch follows the worker count')
opt = parser.parse_args()
logging.basicConfig(level=logging.INFO)
logging.info(opt)

classes = opt.num_classes
num_training_samples = opt.num_training_samples
use_gpu = opt.gpus.strip() != '' and mx.context.num_gpus() > 0
context = [mx.gpu(int(i)) for i in opt.gpus.split(',')] if use_gpu else [mx.cpu()]
kv = mx.kv.create(opt.kvstore)
global_batch_size = opt.batch_size
batch_size = max(1, global_batch_size // kv.num_workers)
epoch_size = int(math.ceil(int(num_training_samples // kv.num_workers) / batch_size))

if opt.lr_mode == 'poly':
    lr_scheduler = mx.lr_scheduler.PolyScheduler(opt.num_epochs * epoch_size, base_lr=opt.lr, pwr=opt.lr_poly_power,
                                                 warmup_steps=opt.warmup_epochs * epoch_size, warmup_begin_lr=opt.warmup_lr)
elif opt.lr_mode == 'step':
    steps = [epoch_size * int(x) for x in opt.lr_decay_epoch.split(',')]
    lr_scheduler = mx.lr_scheduler.MultiFactorScheduler(step=steps, factor=opt.lr_decay, base_lr=opt.lr,
                                                        warmup_steps=opt.warmup_epochs * epoch_size,
                                                        warmup_begin_lr=opt.warmup_lr)
elif opt.lr_mode == 'cosine':
    lr_scheduler = mx.lr_scheduler.CosineScheduler(opt.num_epochs * epoch_size, base_lr=opt.lr,
                                                   warmup_steps=opt.warmup_epochs * epoch_size, warmup_begin_lr=opt.warmup_lr)
else:
    raise ValueError('Invalid lr mode')
if opt.dynamic_minibatch != 'constant_global':
    lr_scheduler = mx.lr_scheduler.DynamicMiniBatchScheduler(lr_scheduler, base_batch=global_batch_size, mode=opt.dynamic_minibatch,
                                                             smooth_steps=epoch_size // 10, momentum=opt.momentum)

kwargs = {'classes': classes}
if opt.last_gamma and 'resnet' in opt.model:
    kwargs['last_gamma'] = True
layout = 'NHWC' if (use_gpu and opt.dtype == 'bfloat16' and opt.mode != 'symbolic') else 'NCHW'
if opt.model == 'resnet50_v1b' and layout == 'NHWC':
    net = vision.get_resnet('v1b', 50, ctx=context[0], dtype=opt.dtype, layout='NHWC', **kwargs)   # fused B200 path
else:
    net = vision.get_model(opt.model, **kwargs)
    if opt.dtype != 'float32':
        net.cast(opt.dtype)

optimizer_params = {'wd': opt.wd, 'momentum': opt.momentum, 'lr_scheduler': lr_scheduler, 'learning_rate': opt.lr}
if opt.dtype != 'float32':
    optimizer_params['multi_precision'] = True


def data_shape():
    s = opt.image_size
    return (s, s, 3) if layout == 'NHWC' else (3, s, s)


def make_iter(batch, num_parts, part_index, train=True):
    if opt.benchmark or not opt.use_rec:
        n = max(batch, (num_training_samples if train else 50000) // num_parts // 64 // batch * batch)
        return mx.io.SyntheticDataIter(classes, (batch,) + data_shape(), max(1, n // batch), dtype=opt.dtype if train else 'float32')
    rec, idx = (opt.rec_train, opt.rec_train_idx) if train else (opt.rec_val, opt.rec_val_idx)
    return mx.io.ImageRecordIter(path_imgrec=rec, path_imgidx=idx, preprocess_threads=opt.num_workers, shuffle=train,
                                 batch_size=batch, data_shape=(3, opt.image_size, opt.image_size),
                                 mean_r=123.68, mean_g=116.779, mean_b=103.939, std_r=58.393, std_g=57.12, std_b=57.375,
                                 rand_mirror=train, rand_crop=train, random_resized_crop=train,
                                 min_random_area=0.08, min_aspect_ratio=3.0 / 4, max_aspect_ratio=4.0 / 3, resize=-1 if train else 256,
                                 num_parts=num_parts, part_index=part_index,
                                 # single-GPU processes decode and augment on the device (nvJPEG + one fused kernel) and
                                 # get the network's input layout / dtype directly; multi-device processes keep host batches
                                 ctx=context[0] if (use_gpu and len(context) == 1) else None, layout=layout,
                                 dtype=opt.dtype if (use_gpu and len(context) == 1 and opt.dtype == 'bfloat16') else 'float32')


class ElasticImageRecordIter(mx.mod.BaseDataIterator):
    """Re-creates the training iterator for the live worker set; called by Module.fit at every epoch boundary."""

    def __init__(self, global_batch):
        self.global_batch = global_batch

    def get_data_iterator(self, kv_):
        b = max(1, self.global_batch // kv_.num_workers)
        logging.info('data iterator for rank %d of %d workers, per-worker batch %d', kv_.rank, kv_.num_workers, b)
        return make_iter(b, kv_.num_workers, kv_.rank, True), None


class ElasticSpeedometer(mx.callback.Speedometer):
    """Speedometer whose batch size follows global_batch / num_workers."""

    def __init__(self, kv_, batch, frequent=50):
        super(ElasticSpeedometer, self).__init__(batch, frequent)
        self._kv = kv_
        self._total = kv_.num_workers * batch

    def __call__(self, param):
        self.batch_size = self._total // self._kv.num_workers
        super(ElasticSpeedometer, self).__call__(param)


def train_symbolic():
    data = mx.sym.var('data')
    if opt.dtype != 'float32':
        data = mx.sym.Cast(data=data, dtype=opt.dtype)
    fused = opt.model.endswith('_v1b') and opt.model.startswith('resnet') and use_gpu and opt.dtype == 'bfloat16'
    if fused:
        # the flagship as ONE graph node running on the B200 kernels (tcgen05 GEMMs, fused BN+ReLU, NHWC inside)
        out = mx.sym.contrib.FusedResNet(data, num_layers=int(opt.model[len('resnet'):-len('_v1b')]), version='v1b',
                                         classes=classes, layout='NCHW', name='resnetv1b')
    else:
        out = net(data)
    if opt.dtype != 'float32':
        out = mx.sym.Cast(data=out, dtype='float32')
    softmax = mx.sym.SoftmaxOutput(out, name='softmax')
    mod = mx.mod.Module(softmax, context=context, data_iterator=ElasticImageRecordIter(global_batch_size))
    save_frequency = opt.save_frequency if opt.save_frequency > 0 else opt.num_epochs
    os.makedirs(opt.save_dir, exist_ok=True)
    train_data = make_iter(batch_size, kv.num_workers, kv.rank, True)
    mod.fit(train_data, eval_data=None, num_epoch=opt.num_epochs, kvstore=kv,
            batch_end_callback=ElasticSpeedometer(kv, batch_size, max(1, opt.log_interval)),
            epoch_end_callback=mx.callback.do_checkpoint(os.path.join(opt.save_dir, 'imagenet-%s' % opt.model), period=save_frequency),
            optimizer='nag', optimizer_params=optimizer_params,
            initializer=mx.init.Xavier(rnd_type='gaussian', factor_type='in', magnitude=2))
    return mod


def train_imperative():
    ctx = context[0]
    if not net.collect_params()._is_initialized() if hasattr(net.collect_params(), '_is_initialized') else True:
        try:
            net.initialize(mx.init.MSRAPrelu(), ctx=ctx)
        except Exception:  # noqa: BLE001  (already initialised by get_resnet)
            pass
    if opt.mode == 'hybrid':
        net.hybridize()
    trainer = gluon.Trainer(net.collect_params(), 'nag', optimizer_params, kvstore=kv)
    if kv.type == 'elastic_device':
        # single-box NVLink path: gradients never leave the GPUs, the scheduler only manages membership; a newcomer
        # adopts the live weights / momentum before its first forward pass
        trainer.init_kvstore()
    L = gluon.loss.SoftmaxCrossEntropyLoss()
    metric = mx.metric.Accuracy()
    begin = int(os.environ.get('EPOCH_BEGIN', '0')) if os.environ.get('NEW_WORKER', '0') == '1' else 0
    for epoch in range(begin, opt.num_epochs):
        if hasattr(kv, '_membership_change_barrier') and not (epoch == begin and begin > 0):
            kv._membership_change_barrier({'EPOCH_BEGIN': str(epoch)})
        b = max(1, global_batch_size // kv.num_workers)
        it = make_iter(b, kv.num_workers, kv.rank, True)
        tic = time.time()
        metric.reset()
        for i, batch in enumerate(it):
            x = batch.data[0].as_in_context(ctx)
            y = batch.label[0].as_in_context(ctx)
            with ag.record():
                out = net(x)
                loss = L(out, y)
            loss.backward()
            trainer.step(global_batch_size)
            metric.update([y], [out])
            if opt.log_interval and (i + 1) % opt.log_interval == 0:
                name, acc = metric.get()
                logging.info('Epoch[%d] Batch [%d]\tSpeed: %f samples/sec (this worker)\t%s=%f', epoch, i + 1,
                             b * opt.log_interval / (time.time() - tic), name, acc)
                tic = time.time()
        if opt.save_frequency and (epoch + 1) % opt.save_frequency == 0:
            os.makedirs(opt.save_dir, exist_ok=True)
            net.save_parameters(os.path.join(opt.save_dir, 'imagenet-%s-%d.params' % (opt.model, epoch)))


if __name__ == '__main__':
    if opt.mode == 'symbolic':
        train_symbolic()
    else:
        train_imperative()
