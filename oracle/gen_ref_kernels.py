"""Oracle (test infrastructure): regenerate the REFERENCE's device kernels for A/B runs.

dtFFT has no ``.cu`` product kernels: ``get_code`` in
``/root/reference/src/dtfft_nvrtc_module.F90:434-583`` builds one CUDA-C template per
kernel kind as a string at run time.  This script does not copy that source: it *reads the
Fortran routine where it lies*, transpiles its control flow (if / select case /
``call code%add``) to Python, executes it for every (kind, ndims, element size) and writes
the emitted CUDA-C -- exactly what NVRTC would be handed -- to ``oracle/_ref/ref_kernels.cu``
(git-ignored; never shipped), together with a thin C launcher restating
``get_kernel_launch_params`` (src/dtfft_kernel_device.F90:227-257).  ``make -C oracle``
compiles it to ``oracle/_ref/libref_kernels.so`` for sm_100a.  Uses: (1) the GPU-side
checker "reference device kernel == our kernel", (2) "the kernel to beat" in bench / kbench.

Runs only where /root/reference exists (this container); the GPU box uses the prebuilt .so.
"""
from __future__ import annotations

import os
import re
import sys

REF = os.environ.get("DTFFT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")


def parse_kernel_constants(src: str):
    consts = {}
    for m in re.finditer(r"::\s*(KERNEL_[A-Z_]+)\s*=\s*kernel_type_t\((-?\d+)\)", src):
        consts[m.group(1)] = int(m.group(2))
    lists = {}
    for m in re.finditer(r"::\s*([A-Z_]+_KERNELS)\(\*\)\s*=\s*\[([^\]]*)\]", src):
        lists[m.group(1)] = [consts[n.strip()] for n in m.group(2).split(",")]
    return consts, lists


def extract_routine(src: str, name: str) -> list[str]:
    lines = src.splitlines()
    start = next(i for i, l in enumerate(lines) if re.match(rf"\s*function {name}\(", l))
    end = next(i for i in range(start, len(lines)) if re.match(rf"\s*end function {name}", lines[i]))
    return lines[start + 1:end]


def tr_expr(e: str) -> str:
    e = e.strip()
    e = re.sub(r"any\(\s*kernel_type\s*==\s*\[([^\]]*)\]\s*\)", r"(kernel_type in (\1,))", e)
    e = e.replace("%val", "")
    e = re.sub(r"\.and\.", " and ", e)
    e = re.sub(r"\.or\.", " or ", e)
    e = re.sub(r"\.not\.", " not ", e)
    e = re.sub(r"\.true\.", "True", e)
    e = re.sub(r"\.false\.", "False", e)
    e = re.sub(r"(\w+)\((\d+):(\d+)\)", lambda m: f"{m.group(1)}[{int(m.group(2)) - 1}:{m.group(3)}]", e)
    e = e.replace("//", "+")
    return e


def transpile(body: list[str]) -> str:
    out, ind = [], 1
    select_stack = []  # (expr, first_case_pending)

    def emit(s):
        out.append("    " * ind + s)

    for raw in body:
        line = raw.split("!!")[0].strip() if '"' not in raw.split("!!")[0] or True else raw.strip()
        if not line or line.startswith("!"):
            continue
        low = line.lower()
        if re.match(r"(character|type\(|integer|logical|class\()", low):
            continue
        if low.startswith("deallocate"):
            continue
        m = re.match(r"allocate\s*\(\s*(\w+)\s*,\s*source\s*=\s*(.*)\)\s*$", line)
        if m:
            emit(f"{m.group(1)} = {m.group(2)}")
            continue
        if low.startswith("internal_error"):
            emit("raise RuntimeError('reference INTERNAL_ERROR')")
            continue
        m = re.match(r"select case\s*\((.*)\)\s*$", line, re.I)
        if m:
            select_stack.append([tr_expr(m.group(1)), True])
            continue
        m = re.match(r"case\s*\((.*)\)\s*$", line, re.I)
        if m:
            expr, first = select_stack[-1]
            if not first:
                ind -= 1
            emit(f"{'if' if first else 'elif'} {expr} in ({tr_expr(m.group(1))},):")
            select_stack[-1][1] = False
            ind += 1
            continue
        if re.match(r"case default", low):
            if not select_stack[-1][1]:
                ind -= 1
            emit("else:" if not select_stack[-1][1] else "if True:")
            select_stack[-1][1] = False
            ind += 1
            continue
        if re.match(r"end\s*select", low):
            if not select_stack.pop()[1]:
                ind -= 1
            continue
        m = re.match(r"else\s*if\s*\((.*)\)\s*then\s*$", line, re.I)
        if m:
            ind -= 1
            emit(f"elif {tr_expr(m.group(1))}:")
            ind += 1
            continue
        m = re.match(r"if\s*\((.*)\)\s*then\s*$", line, re.I)
        if m:
            emit(f"if {tr_expr(m.group(1))}:")
            ind += 1
            continue
        if low == "else":
            ind -= 1
            emit("else:")
            ind += 1
            continue
        if re.match(r"end\s*if", low):
            ind -= 1
            continue
        m = re.match(r"call code%add\((.*)\)\s*$", line)
        if m:
            emit(f"code.append({tr_expr(m.group(1))})")
            continue
        m = re.match(r"(\w+)\s*=\s*(.*)$", line)
        if m:
            emit(f"{m.group(1)} = {tr_expr(m.group(2))}")
            continue
        raise SyntaxError(f"gen_ref_kernels: cannot transpile reference line: {raw!r}")
    return "\n".join(out)


def build_get_code():
    kern_src = open(os.path.join(REF, "src/dtfft_abstract_kernel.F90")).read()
    nvrtc_src = open(os.path.join(REF, "src/dtfft_nvrtc_module.F90")).read()
    consts, lists = parse_kernel_constants(kern_src)
    body = extract_routine(nvrtc_src, "get_code")
    py = "def get_code(kernel_name, ndims, base_storage, kernel_type):\n    code = []\n" + transpile(body) + "\n    return '\\n'.join(code)\n"
    env = dict(consts)
    env.update(FLOAT_STORAGE_SIZE=4, DOUBLE_STORAGE_SIZE=8, DOUBLE_COMPLEX_STORAGE_SIZE=16)
    env["is_transpose_kernel"] = lambda k: k in lists["TRANSPOSE_KERNELS"]
    env["is_unpack_kernel"] = lambda k: k in lists["UNPACK_KERNELS"]
    env["is_pack_kernel"] = lambda k: k in lists["PACK_KERNELS"]
    exec(py, env)
    return env["get_code"], consts


# (tile, rows) instantiations offered to the A/B harness: the candidates that survive the
# reference's filters on B200, PADDING = 1.
CONFIGS = [(16, 4), (16, 8), (16, 16), (32, 4), (32, 8), (32, 16), (32, 32), (64, 4), (64, 8), (64, 16)]
KINDS = ["KERNEL_PERMUTE_FORWARD", "KERNEL_PERMUTE_BACKWARD", "KERNEL_PERMUTE_BACKWARD_START",
         "KERNEL_PERMUTE_BACKWARD_END_PIPELINED", "KERNEL_UNPACK_PIPELINED", "KERNEL_PACK_PIPELINED",
         "KERNEL_PACK_FORWARD", "KERNEL_PACK_BACKWARD"]
CTYPE = {4: "float", 8: "double", 16: "double2"}


def main():
    get_code, consts = build_get_code()
    os.makedirs(OUT_DIR, exist_ok=True)
    parts = ["// GENERATED by oracle/gen_ref_kernels.py from the reference's get_code(); do not commit.",
             "#include <cuda_runtime.h>", ""]
    table = []
    for kind in KINDS:
        kt = consts[kind]
        for ndims in (2, 3):
            if ndims == 2 and kind in ("KERNEL_PERMUTE_BACKWARD", "KERNEL_PERMUTE_BACKWARD_START",
                                       "KERNEL_PERMUTE_BACKWARD_END_PIPELINED", "KERNEL_PACK_BACKWARD"):
                continue
            for es in (4, 8, 16):
                name = f"ref_k{kt}_d{ndims}_s{es}"
                parts.append(get_code(name, ndims, es, kt))
                parts.append("")
                table.append((kt, ndims, es, name))
    # launcher: restates get_kernel_launch_params / get_kernel_args (kernel_device.F90:188-257)
    packers = {consts[k] for k in ("KERNEL_PERMUTE_BACKWARD_END_PIPELINED", "KERNEL_UNPACK_PIPELINED",
                                   "KERNEL_PACK_PIPELINED", "KERNEL_PACK_FORWARD", "KERNEL_PACK_BACKWARD")}
    tile2 = {consts[k] for k in ("KERNEL_PERMUTE_FORWARD", "KERNEL_PACK_FORWARD", "KERNEL_PERMUTE_BACKWARD_END_PIPELINED",
                                 "KERNEL_UNPACK_PIPELINED", "KERNEL_PACK_PIPELINED")}
    L = ['extern "C" int ref_kernel_launch(int kernel_type, int ndims, int es, int tile, int rows, void* out, const void* in,',
         '        int nx, int ny, int nz, int nxx, int nyy, int nzz, int din, int dout, void* stream) {',
         '    cudaStream_t s = (cudaStream_t)stream;',
         f'    const bool packer = {" || ".join(f"kernel_type == {k}" for k in sorted(packers))};',
         f'    const bool tile_dim2 = {" || ".join(f"kernel_type == {k}" for k in sorted(tile2))};',
         '    int d1 = packer ? nxx : nx, d2 = packer ? nyy : ny, d3 = packer ? nzz : nz;',
         '    dim3 threads(tile, rows, 1);',
         '    dim3 blocks((d1 + tile - 1) / tile, ((tile_dim2 ? d2 : d3) + tile - 1) / tile, ndims == 2 ? 1 : (tile_dim2 ? d3 : d2));',
         '    if (blocks.x == 0 || blocks.y == 0 || blocks.z == 0) return 0;']
    for (kt, ndims, es, name) in table:
        for (t, r) in CONFIGS:
            if t * (t + 1) * es >= 0.9 * 48 * 1024 or t * r > 1024 or t * r < 64 or t < r:
                continue
            args = "(%s*)out, (const %s*)in, nx, ny, nz" % (CTYPE[es], CTYPE[es])
            if kt in packers:
                args += ", nxx, nyy, nzz, din, dout"
            L.append(f'    if (kernel_type == {kt} && ndims == {ndims} && es == {es} && tile == {t} && rows == {r}) '
                     f'{{ {name}<{t}, {r}, 1><<<blocks, threads, 0, s>>>({args}); return (int)cudaGetLastError(); }}')
    L += ['    return -1;', '}', '']
    cfgs = ", ".join(f"{t}, {r}" for t, r in CONFIGS)
    L += [f'extern "C" int ref_kernel_configs(int* out, int max_pairs) {{ static const int c[] = {{{cfgs}}}; '
          f'int n = {len(CONFIGS)}; if (n > max_pairs) n = max_pairs; for (int i = 0; i < 2 * n; ++i) out[i] = c[i]; return n; }}', '']
    with open(os.path.join(OUT_DIR, "ref_kernels.cu"), "w") as f:
        f.write("\n".join(parts + L))
    print(f"wrote {os.path.join(OUT_DIR, 'ref_kernels.cu')}: {len(table)} reference kernels")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("gen_ref_kernels: reference tree not present, nothing to do")
        sys.exit(0)
    main()
