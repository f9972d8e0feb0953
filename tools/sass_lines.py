"""Static SASS instruction count per CUDA source line of one kernel (no GPU needed).

    python tools/sass_lines.py build/csrc/spec_kernels.o spec_frames_kernelILi1024ELb0 [top]

Uses `cuobjdump -xelf` + `nvdisasm -g`; the object must be compiled with -lineinfo.  Instructions
are attributed to the innermost source line nvdisasm reports (inlined callees count for themselves).
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
        text = "".join(subprocess.run(["nvdisasm", "-g", c], capture_output=True, text=True).stdout for c in cubins)
    cur_fun, cur = None, "?"
    cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
    for line in text.splitlines():
        m = re.match(r"\.text\.(\S+):", line)
        if m:
            cur_fun = m.group(1)
            continue
        if cur_fun is None or pat not in cur_fun:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = m.group(1).split("/")[-1] + ":" + m.group(2)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m:
            cnt[cur] += 1
            ops[cur][m.group(1)] += 1
    print("total static instructions", sum(cnt.values()))
    for k, v in cnt.most_common(top):
        print(f"{v:5d} {k:28s} {dict(ops[k].most_common(5))}")


if __name__ == "__main__":
    main()
