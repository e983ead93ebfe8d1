"""Per-kernel comparison of the SASS encodings of two object files / shared libraries (cuobjdump -sass).

    python scripts/compare_sass.py old.o new.o

Used after source clean-ups that must not change generated code (dead template variants removed, helpers moved): every
kernel present in both files is compared by its 64-bit instruction encodings, keyed by demangled name (the anonymous
namespace hash in the mangled name depends on the file path, so mangled names cannot be used).  Exit code = number
of kernels whose code differs.  Template-parameter renames need a mapping by hand (see git history of this file's
first use: run_update_kernel lost its last parameter, dot_interact_fwd_kernel its third)."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    code, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            code[name] = []
        elif name:
            code[name] += re.findall(r"/\* (0x[0-9a-f]{16}) \*/", line)
    names = list(code)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return {d: code[n] for n, d in zip(names, dem)}


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in new if k in old and old[k] == new[k]]
    diff = [k for k in new if k in old and old[k] != new[k]]
    print(f"{len(same)} kernels bit-identical, {len(diff)} different, {len([k for k in new if k not in old])} only in "
          f"{sys.argv[2]}, {len([k for k in old if k not in new])} only in {sys.argv[1]}")
    for k in diff:
        print("  DIFFERENT:", k[:160])
    return len(diff)


if __name__ == "__main__":
    sys.exit(main())
