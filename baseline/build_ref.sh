#!/bin/bash
# Build attempt of the UNMODIFIED reference (MXNet 1.3.1 fork, /root/reference) for this box: CUDA 12.9, sm_100,
# no cuDNN-7 API (USE_CUDNN=0: the reference's own non-cuDNN operator path), no OpenCV, no network.
# The source tree is copied to $WORK (the reference mount is read-only); every step logs to baseline/ref_build.log.
# Result on success: baseline/_ref/mxnet (python package, unmodified) + baseline/_ref/mxnet/libmxnet.so.
#
#   bash baseline/build_ref.sh [jobs]
#
# Toolchain shims (none touches a reference source line that executes in the benchmark):
#   * make/config.mk is git-ignored upstream -> we write one (the documented knobs only)
#   * 3rdparty/cub (2018) shadows the CUDA-12 toolkit's own cub/thrust -> the directory is emptied in the COPY
#   * cblas.h is not in the image -> a 40-line declaration header for the OpenBLAS that ships with opencv-python
#   * patches under baseline/patches/*.patch are compile fixes for CUDA 12 / gcc 13, applied to the COPY and listed in
#     DESIGN.md (01: `cuda::Reduce1D` -> `mshadow::cuda::Reduce1D`, ambiguous with libcu++'s global ::cuda namespace)
#     (02: `#include <unistd.h>` for the getpid() call the fork added to kvstore_local.h)
#     (04: explicit instantiations of three sparse elementwise member templates that other TUs reference)
#     (03: one more `cuda::Reduce1D`; random/sampler.h's file-scope `using namespace mshadow` moved inside
#          namespace mxnet::op -- at file scope it makes `cuda::` ambiguous in nvcc's generated host stubs)
#   * `--pre-include cub/cub.cuh` (nvcc) parses the toolkit's cub before any `using namespace mshadow`
set -u
JOBS=${1:-6}
REPO=$(cd "$(dirname "$0")/.." && pwd)
WORK=${MXREF_WORK:-/tmp/mxnet_ref}
LOG=$REPO/baseline/ref_build.log
exec > >(tee -a "$LOG") 2>&1
echo "=== build_ref.sh $(date -u +%FT%TZ) jobs=$JOBS work=$WORK"

if [ ! -d "$WORK/src" ]; then
  mkdir -p "$WORK"
  for d in 3rdparty/dlpack 3rdparty/dmlc-core 3rdparty/mshadow 3rdparty/ps-lite 3rdparty/tvm include src python make \
           mkldnn.mk Makefile plugin example/image-classification tools; do
    mkdir -p "$WORK/$(dirname $d)"
    cp -r /root/reference/$d "$WORK/$d"
  done
  mkdir -p "$WORK/3rdparty/cub"     # emptied on purpose: use the toolkit's cub
fi
# pieces of the upstream tree that are ABSENT from the reference snapshot (git-ignored / un-fetched submodules): the code
# that needs them cannot be built and is left out of the copy -- none of it is on the image-classification path
#   3rdparty/ctc_include (warp-ctc)        -> operator CTCLoss
#   3rdparty/dmlc-core/src/data/ (parsers) -> LibSVMIter / CSVIter and dmlc-core's data.o
rm -f "$WORK"/src/operator/nn/ctc_loss.cc "$WORK"/src/operator/nn/ctc_loss.cu "$WORK"/src/operator/contrib/ctc_loss.cc "$WORK"/src/operator/contrib/ctc_loss.cu
rm -f "$WORK"/src/io/iter_libsvm.cc "$WORK"/src/io/iter_csv.cc
#   quantized_conv.cu is a cuDNN-only operator (its whole body is `#if MXNET_USE_CUDNN`); with USE_CUDNN=0 it only
#   registers a LOG(FATAL) stub, and its host stub does not compile against the CUDA-12 libcu++ -> left out
rm -f "$WORK"/src/operator/quantization/quantized_conv.cu
sed -i 's/ data.o recordio.o/ recordio.o/' "$WORK/3rdparty/dmlc-core/Makefile"
cd "$WORK" || exit 1

OB=$(ls /opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs/libopenblasp-*.so | head -1)
mkdir -p compat/lib
ln -sf "$OB" compat/lib/libopenblas.so
cat > compat/cblas.h <<'EOF'
// minimal CBLAS declarations (the image ships OpenBLAS binaries without headers)
#ifndef COMPAT_CBLAS_H_
#define COMPAT_CBLAS_H_
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
typedef enum CBLAS_ORDER CBLAS_LAYOUT;
void cblas_sgemm(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, enum CBLAS_TRANSPOSE, int, int, int, float, const float*, int,
                 const float*, int, float, float*, int);
void cblas_dgemm(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, enum CBLAS_TRANSPOSE, int, int, int, double, const double*, int,
                 const double*, int, double, double*, int);
void cblas_sgemv(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int, int, float, const float*, int, const float*, int, float,
                 float*, int);
void cblas_dgemv(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int, int, double, const double*, int, const double*, int, double,
                 double*, int);
void cblas_sger(enum CBLAS_ORDER, int, int, float, const float*, int, const float*, int, float*, int);
void cblas_dger(enum CBLAS_ORDER, int, int, double, const double*, int, const double*, int, double*, int);
float cblas_sdot(int, const float*, int, const float*, int);
double cblas_ddot(int, const double*, int, const double*, int);
void cblas_saxpy(int, float, const float*, int, float*, int);
void cblas_daxpy(int, double, const double*, int, double*, int);
void cblas_strsm(enum CBLAS_ORDER, enum CBLAS_SIDE, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, enum CBLAS_DIAG, int, int,
                 float, const float*, int, float*, int);
void cblas_dtrsm(enum CBLAS_ORDER, enum CBLAS_SIDE, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, enum CBLAS_DIAG, int, int,
                 double, const double*, int, double*, int);
void cblas_strmm(enum CBLAS_ORDER, enum CBLAS_SIDE, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, enum CBLAS_DIAG, int, int,
                 float, const float*, int, float*, int);
void cblas_dtrmm(enum CBLAS_ORDER, enum CBLAS_SIDE, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, enum CBLAS_DIAG, int, int,
                 double, const double*, int, double*, int);
void cblas_ssyrk(enum CBLAS_ORDER, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, int, int, float, const float*, int, float,
                 float*, int);
void cblas_dsyrk(enum CBLAS_ORDER, enum CBLAS_UPLO, enum CBLAS_TRANSPOSE, int, int, double, const double*, int, double,
                 double*, int);
#ifdef __cplusplus
}
#endif
#endif
EOF

cat > make/config.mk <<EOF
export CC = /usr/bin/gcc
export CXX = /usr/bin/g++
export NVCC = /usr/local/cuda/bin/nvcc
DEV = 0
DEBUG = 0
USE_SIGNAL_HANDLER = 0
ADD_LDFLAGS = -L$WORK/compat/lib -Wl,-rpath,$(dirname "$OB") -L/usr/local/cuda/lib64
ADD_CFLAGS = -I$WORK/compat -Wno-deprecated-declarations -Wno-unused -Wno-class-memaccess -Wno-maybe-uninitialized -fpermissive
USE_CUDA = 1
USE_CUDA_PATH = /usr/local/cuda
ENABLE_CUDA_RTC = 0
USE_CUDNN = 0
USE_NCCL = 1
USE_NCCL_PATH = NONE
USE_OPENCV = 0
USE_LIBJPEG_TURBO = 0
USE_OPENMP = 1
USE_MKLDNN = 0
USE_NNPACK = 0
USE_BLAS = openblas
USE_LAPACK = 0
USE_INTEL_PATH = NONE
USE_STATIC_MKL = NONE
USE_SSE = 1
USE_F16C = 0
USE_DIST_KVSTORE = 0
USE_HDFS = 0
USE_S3 = 0
USE_OPERATOR_TUNING = 1
USE_GPERFTOOLS = 0
USE_JEMALLOC = 0
USE_THREADED_ENGINE = 1
USE_CPP_PACKAGE = 0
USE_TENSORRT = 0
CUDA_ARCH = -gencode arch=compute_100,code=sm_100
MXNET_PLUGINS =
EXTRA_OPERATORS =
USE_PROFILER = 1
EOF

if ls "$REPO"/baseline/patches/*.patch >/dev/null 2>&1; then
  for p in "$REPO"/baseline/patches/*.patch; do
    if [ ! -f ".applied_$(basename $p)" ]; then
      echo "--- applying $(basename $p)"
      patch -p1 --forward < "$p" && touch ".applied_$(basename $p)"
    fi
  done
fi

echo "--- make -k -j$JOBS lib/libmxnet.so"
make -k -j"$JOBS" lib/libmxnet.so NVCCFLAGS="-Wno-deprecated-gpu-targets --pre-include cub/cub.cuh -Xcompiler -fpermissive --expt-relaxed-constexpr -diag-suppress 20012,20014,186,611,997,2810,177" 2>&1 | grep -v "^ALL_DEP is" | tail -n 4000
RC=${PIPESTATUS[0]}
echo "--- make exit $RC"
if [ -f lib/libmxnet.so ]; then
  rm -rf "$REPO/baseline/_ref"
  mkdir -p "$REPO/baseline/_ref"
  cp -r python/mxnet "$REPO/baseline/_ref/mxnet"
  cp lib/libmxnet.so "$REPO/baseline/_ref/mxnet/libmxnet.so"
  cp "$OB" "$REPO/baseline/_ref/mxnet/"
  cp "$(dirname "$OB")"/libgfortran-*.so* "$(dirname "$OB")"/libquadmath-*.so* "$REPO/baseline/_ref/mxnet/"   # OpenBLAS runtime deps
  mkdir -p "$REPO/baseline/_ref/example"
  cp -r example/image-classification "$REPO/baseline/_ref/example/"
  echo "=== SUCCESS: baseline/_ref/mxnet/libmxnet.so"
else
  echo "=== FAILED: no lib/libmxnet.so (see errors above)"
fi
