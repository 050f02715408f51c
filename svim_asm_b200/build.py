"""Build libsvimasm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsvimasm_b200.so")
SOURCES = ["capi.cu", "cigar_scan.cu", "segment_walk.cu", "pair.cu", "edit_distance.cu", "wfa.cu", "seqpool.cu", "exchange.cu", "bam_device.cu", "fasta_device.cu", "vcf_device.cu", "file_upload.cu", "bam_ingest.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--threads", "4"]


def _stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "svimasm_b200.h")]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force=False, verbose=False, extra_flags=(), out=None):
    """Compile every CUDA/C++ source of the package into one shared library.  Returns its path.
    `extra_flags` / `out` build tuning variants (e.g. -DK2_CHUNKS_PER_WARP=8) next to the default library."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objs = []
    build_dir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, proc in procs:
        text, _ = proc.communicate()
        if verbose or proc.returncode:
            sys.stderr.write(text)
        if proc.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    target = LIB if out is None else out
    link = [nvcc, "-shared", "-o", target] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lz", "-lpthread"]
    subprocess.check_call(link)
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
