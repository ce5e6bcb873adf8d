#!/bin/bash
# Elastic ResNet-50 v1b recipe of the reference (example/dynamic-training/run.sh:16) on mxb200.
# $DEEPLEARNING_WORKERS_COUNT workers listed in $DEEPLEARNING_WORKERS_PATH; edit the generated host_worker file (or let the
# EC2 manager thread of tools/launch.py do it) to add / remove workers at the next epoch boundary.
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
python $ROOT/tools/launch.py -n ${DEEPLEARNING_WORKERS_COUNT:-2} ${DEEPLEARNING_WORKERS_PATH:+-H $DEEPLEARNING_WORKERS_PATH} \
  --launcher ${LAUNCHER:-local} --elastic-training-enabled True \
  python $ROOT/examples/dynamic-training/train_resnet.py --use-rec --batch-size 4096 --dtype bfloat16 --num-data-workers 40 \
  --num-epochs 90 --gpus 0,1,2,3,4,5,6,7 --lr 1.6 --lr-mode poly --warmup-epochs 5 --last-gamma --mode symbolic \
  --model resnet50_v1b --kvstore dist_sync_device --rec-train ${TRAIN_REC:-train.rec} --rec-train-idx ${TRAIN_IDX:-train.idx} \
  --rec-val ${VAL_REC:-val.rec} --rec-val-idx ${VAL_IDX:-val.idx} --save-frequency 1000 --warmup-lr 0.001 "$@"
