"""Per-kernel fingerprint of the SASS in an object file or shared library: registers, stack, and an md5 of the
instruction stream (addresses and encodings stripped).  Used to prove that a change which adds kernels or template
parameters leaves the existing (measured) kernels bit-identical:
    python tools/sass_fingerprint.py galaxy_b200/csrc/*.o > before.json ; <change> ; ... > after.json ; diff
"""
import hashlib
import json
import re
import subprocess
import sys


def fingerprint(path):
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    usage, name = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            usage[name] = dict(re.findall(r"(\w+):(\d+)", line))
            name = None
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, cur, h, n = {}, None, None, 0

    def close():
        if cur:
            out[cur] = dict(md5=h.hexdigest(), instructions=n,
                            **{k: int(v) for k, v in usage.get(cur, {}).items() if k in ("REG", "STACK", "SHARED")})
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            close()
            cur, h, n = m.group(1), hashlib.md5(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*", line)
        if m and cur:
            # parameter offsets in the constant bank move when a parameter struct grows at its end: not a code change
            h.update(re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0][*]", m.group(1)).encode())
            n += 1
    close()
    return out


if __name__ == "__main__":
    all_ = {}
    for p in sys.argv[1:]:
        all_.update(fingerprint(p))
    json.dump(all_, sys.stdout, indent=0, sort_keys=True)
