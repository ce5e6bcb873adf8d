"""Runtime compilation of user CUDA kernels (reference: python/mxnet/rtc.py CudaModule, src/common/rtc.cc NVRTC).
Implemented with nvcc -> cubin -> cuModuleLoad through cuda-python when available."""


class CudaModule(object):
    def __init__(self, source, options=(), exports=()):
        import subprocess
        import tempfile
        import os
        self._dir = tempfile.mkdtemp(prefix="mxb200_rtc_")
        src = os.path.join(self._dir, "k.cu")
        with open(src, "w") as f:
            f.write(source)
        self._cubin = os.path.join(self._dir, "k.cubin")
        subprocess.check_call(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-o", self._cubin, src] + list(options))
        self._mod = None
        self.exports = list(exports)

    def _load(self):
        if self._mod is None:
            from cuda import cuda
            import torch
            torch.cuda.init()
            with open(self._cubin, "rb") as f:
                data = f.read()
            err, self._mod = cuda.cuModuleLoadData(data)
            if int(err) != 0:
                raise RuntimeError("cuModuleLoadData failed: %s" % err)
        return self._mod

    def get_kernel(self, name, signature=None):
        from cuda import cuda
        err, fn = cuda.cuModuleGetFunction(self._load(), name.encode())
        if int(err) != 0:
            raise RuntimeError("kernel %s not found" % name)
        return CudaKernel(fn, name)


class CudaKernel(object):
    def __init__(self, fn, name):
        self._fn, self._name = fn, name

    def launch(self, args, ctx, grid_dims, block_dims, shared_mem=0):
        import ctypes
        import torch
        from cuda import cuda
        vals = []
        for a in args:
            if hasattr(a, "_t"):
                vals.append(ctypes.c_void_p(a._t.data_ptr()))
            elif isinstance(a, int):
                vals.append(ctypes.c_int(a))
            else:
                vals.append(ctypes.c_float(a))
        ptrs = (ctypes.c_void_p * len(vals))(*[ctypes.addressof(v) for v in vals])
        stream = torch.cuda.current_stream().cuda_stream
        err, = cuda.cuLaunchKernel(self._fn, *grid_dims, *block_dims, shared_mem, stream, ctypes.addressof(ptrs), 0)
        if int(err) != 0:
            raise RuntimeError("cuLaunchKernel failed: %s" % err)
