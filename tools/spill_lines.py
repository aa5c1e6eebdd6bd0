"""List local-memory (spill) instructions of a function in an object file by CUDA source line.
    python tools/spill_lines.py <file.o|.so> <function-substring>"""
import collections, os, re, subprocess, sys, tempfile
obj, kname = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        lines = out.split("\n")
        for si, l in enumerate(lines):
            if ".section\t.text." in l and kname in l:
                end = next((i for i in range(si + 1, len(lines)) if lines[i].startswith("//---------------------")), len(lines))
                cur = ("?", 0); cnt = collections.Counter(); n = 0; fn = "kernel"
                for l2 in lines[si:end]:
                    mf = re.match(r"\s*(\.?_Z\w+):", l2)
                    if mf and "$" not in mf.group(1): fn = re.sub(r"^_ZN\d*\w*?3gsf\d+", "", mf.group(1))[:28]
                    m = re.search(r'//## File "([^"]+)", line (\d+)', l2)
                    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
                    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l2):
                        n += 1
                        mm = re.search(r"\b(LDL|STL)[.\w]*", l2)
                        if mm: cnt[(fn, cur, mm.group(1))] += 1
                print(l.strip()[:140], "instructions", n)
                for k, v in sorted(cnt.items()): print("   ", k, v)
