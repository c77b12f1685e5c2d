"""TEST INFRASTRUCTURE: builds the product's CUDA kernels for the SIMT emulator (tests/simt/simt.h).

The kernel sources are COPIED from ngsf-hmm_b200/csrc into a scratch directory and rewritten mechanically - every
rewrite asserts that it found what it expected, so a change of the kernels that the rewrites no longer cover fails
loudly instead of testing stale text:

  * `kernel<<<grid, block, smem, stream>>>(args);`  ->  `simt::launch(grid, block, smem, [&]() { kernel(args); });`
  * `extern __shared__ __align__(n) T name[];`       ->  `T *name = (T *) simt::dyn_smem();`
  * the warp-level helpers of nfh_device.cuh / nfh_math.cuh that are guarded by `__CUDACC__` are switched on
    (shuffles and `threadIdx` exist in the emulator); the inline-PTX reciprocal seed keeps its host stand-in
  * `__shared__ alignas(n)` -> `alignas(n) __shared__` (`__shared__` is `static` here and ISO C++ wants that order)
  * nfh_tma.cuh is not used: simt.h provides synchronous bulk copies and mbarriers of the same names
  * nfh_freq.cu: the named barrier of the team kernel (`bar.sync id, n`, inline PTX) -> simt::named_barrier; the FP64
    probe of bench.py is cut; the tensor maps are encoded by the product's own freq_tensor_maps() through an emulated
    cuTensorMapEncodeTiled and honoured by the emulator's tma_load_2d (box, out-of-bounds zero fill, swizzle)
  * nfh_estep.cu: the inline PTX of the opt-in single-launch variant (`estep_fused`: acquire / release on cross-CTA
    flags, bulk-store group wait) becomes plain memory operations - CTAs run one after the other here

Nothing else changes: the kernels' bodies, their launchers and their launch geometry are the product's text.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ngsf-hmm_b200", "csrc")
SIMT = os.path.join(ROOT, "tests", "simt")


def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(src):
    """kernel<<<g, b, s, st>>>(args);  ->  simt::launch(...)"""
    out, pos, n = "", 0, 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            break
        m = re.search(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*$", src[pos:i])
        assert m, "no kernel name before <<<"
        name_start = pos + m.start(1)
        j = src.index(">>>", i)
        cfg = _split_top_level(src[i + 3:j])
        assert len(cfg) == 4, cfg
        k = src.index("(", j)
        depth, e = 0, k
        for e in range(k, len(src)):
            depth += src[e] == "("
            depth -= src[e] == ")"
            if depth == 0:
                break
        assert src[e + 1] == ";", src[e:e + 20]
        args = src[k + 1:e]
        out += src[pos:name_start]
        out += (f"simt::launch(dim3({cfg[0]}), dim3({cfg[1]}), (size_t) ({cfg[2]}), "
                f"[&]() {{ {m.group(1)}({args}); }});")
        pos = e + 2
        n += 1
    return out + src[pos:], n


def rewrite_dyn_smem(src):
    pat = re.compile(r"extern __shared__ __align__\(\d+\) (unsigned char|double) (\w+)\[\];")
    return pat.subn(lambda m: f"{m.group(1)} *{m.group(2)} = reinterpret_cast<{m.group(1)} *>(simt::dyn_smem());", src)


def _cut(src, begin, end, what):
    a = src.index(begin)
    b = src.index(end, a)
    assert a < b, what
    return src[:a] + src[b:]


def transform(name, src):
    if name == "nfh_math.cuh":
        old = "#if defined(__CUDACC__)\n__device__ __forceinline__ void load_exp_table"
        assert src.count(old) == 1
        src = src.replace(old, "#if 1\n__device__ __forceinline__ void load_exp_table")
    if name == "nfh_device.cuh":
        old = "#if defined(__CUDACC__)   // warp-level pieces: device only"
        assert src.count(old) == 1
        src = src.replace(old, "#if 1")
    if name == "nfh_estep.cu":
        # the opt-in single-launch variant (estep_fused): its inline PTX - acquire loads / release stores / a release
        # reduction on the cross-CTA flags, and the bulk-store group wait - becomes plain memory operations (CTAs run
        # one after the other here, so the first CTA takes every ticket and never waits for another one)
        for old, new in [
                ('asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");', "v = *p;"),
                ('asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");', "v = *p;"),
                ('asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");', "*p = v;"),
                ('asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");', "*p += 1;"),
                ('asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");', "simt::flush_stores();")]:
            assert src.count(old) == 1, old
            src = src.replace(old, new)
        assert "asm" not in re.sub(r"//[^\n]*", "", src)
    if name == "nfh_freq.cu":
        old = 'asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(n_threads) : "memory");'
        assert src.count(old) == 1
        src = src.replace(old, "simt::named_barrier(team + 1, n_threads);")
        # the FP64 probe of bench.py: events and device allocation, nothing to check on the host
        a = src.index("double launch_fp64_probe(")
        src = src[:a] + "}  // namespace nfh\n"
        assert "asm" not in re.sub(r"//[^\n]*", "", src)
    if name.endswith(".cu"):
        src = re.sub(r"__shared__ alignas\((\d+)\)", r"alignas(\1) __shared__", src)   # ISO order for `static`
        src, n_launch = rewrite_launches(src)
        src, n_smem = rewrite_dyn_smem(src)
        expect = {"nfh_estep.cu": (4, 3), "nfh_lkl.cu": (2, 1), "nfh_viterbi.cu": (5, 2), "nfh_freq.cu": (11, 3)}.get(name)
        if expect:
            assert (n_launch, n_smem) == expect, (name, n_launch, n_smem)
    return src


def build(scratch, flags=("-ffp-contract=off",), extra_sources=(), extra_objects=(), name="libsimt_kernels.so"):
    """Rewrites the kernel sources into `scratch`, compiles the harness (plus `extra_sources`, linked with
    `extra_objects`); returns the shared library's path."""
    os.makedirs(scratch, exist_ok=True)
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh", ".h")) or f == "nfh_tma.cuh":
            continue
        text = open(os.path.join(CSRC, f)).read()
        open(os.path.join(scratch, f), "w").write(transform(f, text))
    so = os.path.join(scratch, name)
    # translation units side by side: the frequency kernels are most of the compile time (hundreds of instantiations)
    # and stay at -O0; the scan kernels, where the long emulated runs spend their time, get -O1
    common = ["-std=c++17", "-fPIC", "-c", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable",
              "-Wno-unused-but-set-variable", "-Wno-psabi"] + list(flags) + ["-I", scratch, "-I", SIMT, "-I",
                                                                            os.path.join(ROOT, "include")]
    units = [(os.path.join(ROOT, "tests", "simt_kernels_host.cpp"), "-O1"),
             (os.path.join(ROOT, "tests", "simt_freq_host.cpp"), "-O0"), (os.path.join(SIMT, "simt.cpp"), "-O1")]
    units += [(src_file, "-O1") for src_file in extra_sources]
    procs = []
    for src_file, opt in units:
        obj = os.path.join(scratch, os.path.basename(src_file) + ".o")
        procs.append((subprocess.Popen(["g++", opt] + common + ["-o", obj, src_file]), src_file, obj))
    for pr, src_file, _ in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, src_file)
    subprocess.check_call(["g++", "-shared", "-o", so] + [o for _, _, o in procs] + list(extra_objects) +
                          ["-Wl,--no-undefined"])
    return so


def build_whole_product(scratch, opt="-O1", freq_opt="-O1"):
    """The whole product for the emulator, in `scratch`: libngsfhmm_b200.so (kernels + launchers + nfh_ctx.cu),
    libngsfhmm_host.so and the ngsF-HMM binary from the product's host sources, linked as host/Makefile links them."""
    os.makedirs(scratch, exist_ok=True)
    src = os.path.join(scratch, "src")
    os.makedirs(src, exist_ok=True)
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")) and f != "nfh_tma.cuh":
            open(os.path.join(src, f), "w").write(transform(f, open(os.path.join(CSRC, f)).read()))
    inc = os.path.join(ROOT, "include")
    cuda_so = os.path.join(scratch, "libngsfhmm_b200.so")
    host = os.path.join(ROOT, "ngsf-hmm_b200", "host")
    # every translation unit at once: the two units of the emulated library (the frequency kernels - hundreds of
    # instantiations - are 47 s at -O1, 15 s at -O0; everything else 4 s), the emulator, and the product's host and
    # command-line sources with the flags of host/Makefile
    common = ["-std=c++17", "-fPIC", "-c", "-w", "-ffp-contract=off", "-I", src, "-I", SIMT, "-I", inc]
    hflags = ["-O2", "-std=c++17", "-fPIC", "-c", "-Wall", "-Wextra", "-ffp-contract=off", "-I", inc, "-I", host]
    units = [("libngsfhmm_b200_emulated.cpp", opt), ("libngsfhmm_b200_emulated_freq.cpp", freq_opt), ("simt.cpp", opt)]
    lib_src = ["lbfgsb.cpp", "bfgs_driver.cpp", "host_api.cpp", "group.cpp"]
    cli_src = sorted(f for f in os.listdir(os.path.join(host, "cli")) if f.endswith(".cpp"))
    procs = [(subprocess.Popen(["g++", o] + common + ["-o", os.path.join(scratch, u + ".o"), os.path.join(SIMT, u)]), u)
             for u, o in units]
    procs += [(subprocess.Popen(["g++"] + hflags + ["-o", os.path.join(scratch, "host_" + f + ".o"), os.path.join(host, f)]), f)
              for f in lib_src]
    procs += [(subprocess.Popen(["g++"] + hflags + ["-o", os.path.join(scratch, "cli_" + f + ".o"),
                                                    os.path.join(host, "cli", f)]), f) for f in cli_src]
    for pr, u in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, u)
    subprocess.check_call(["g++", "-shared", "-o", cuda_so] + [os.path.join(scratch, u + ".o") for u, _ in units] +
                          ["-lpthread"])
    host_so = os.path.join(scratch, "libngsfhmm_host.so")
    subprocess.check_call(["g++", "-shared", "-o", host_so] + [os.path.join(scratch, "host_" + f + ".o") for f in lib_src] +
                          ["-L", scratch, "-lngsfhmm_b200", "-lpthread", "-Wl,-rpath,$ORIGIN"])
    cli = os.path.join(scratch, "ngsF-HMM")
    subprocess.check_call(["g++", "-o", cli] + [os.path.join(scratch, "cli_" + f + ".o") for f in cli_src] +
                          ["-L", scratch, "-lngsfhmm_host", "-lngsfhmm_b200", "-lz", "-lpthread", "-Wl,-rpath,$ORIGIN"])
    # the reference's own objects with iter_EM / viterbi overridden through the C ABI
    # (oracle/Makefile, target `patched`), linked against the emulated library instead of the real one
    obj = os.path.join(ROOT, "oracle", "_ref", "obj")
    need = ["EM_weak.o", "HMM_weak.o", "b200_patch.o", "ngsF-HMM.o"]
    if all(os.path.exists(os.path.join(obj, f)) for f in need):
        rest = sorted(f for f in os.listdir(obj) if f.endswith(".o") and f not in need + ["EM.o", "HMM.o", "ref_harness.o"])
        subprocess.check_call(["g++", "-O3"] + [os.path.join(obj, f) for f in rest + need] +
                              ["-L", scratch, "-lngsfhmm_host", "-lngsfhmm_b200", "-lz", "-lpthread", "-Wl,-rpath,$ORIGIN",
                               "-o", os.path.join(scratch, "ngsF-HMM_b200patch")])
    return cuda_so, host_so, cli
