"""Extract the numeric outputs MATLAB embedded in the reference live script
`utils/One_code.mlx` (a zip: code in matlab/document.xml, values in matlab/output.xml)
into tests/golden/mlx_one_code.json.  Run in the build container only:

    python tests/golden/make_mlx_golden.py [/root/reference]

Values are MATLAB `format short` prints (4-5 significant digits); a leading
"1.0e-03 *"-style scale factor line is applied.  The JSON keeps, per variable, the
live-script line number it was printed from (the same name can be printed twice).
"""
import json
import os
import re
import sys
import zipfile


def parse_value(txt):
    rows, scale, blocks, cur = [], 1.0, [], []
    for line in txt.splitlines():
        s = line.strip()
        if not s:
            continue
        m = re.match(r"^1\.0e([+-]\d+)\s*\*$", s)
        if m:
            scale = 10.0 ** int(m.group(1))
            continue
        if s.startswith("Columns") or s.startswith("Column"):
            if cur:
                blocks.append(cur)
                cur = []
            continue
        try:
            cur.append([float(t) for t in s.split()])
        except ValueError:
            return None
    if cur:
        blocks.append(cur)
    if not blocks:
        return None
    nrow = len(blocks[0])
    rows = [[] for _ in range(nrow)]
    for b in blocks:
        if len(b) != nrow:
            return None
        for i, r in enumerate(b):
            rows[i].extend(v * scale for v in r)
    return rows


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    z = zipfile.ZipFile(os.path.join(ref, "utils", "One_code.mlx"))
    out = z.read("matlab/output.xml").decode()
    items = []
    for el in re.findall(r"<element><type>(?:matrix|variable)</type>(.*?)</lineNumbers></element>", out, flags=re.S):
        name = re.search(r"<name>(.*?)</name>", el, flags=re.S).group(1)
        val = re.search(r"<value>(.*?)</value>", el, flags=re.S).group(1)
        rows_decl = re.search(r"<rows>(\d+)</rows>", el)
        cols_decl = re.search(r"<columns>(\d+)</columns>", el)
        line = re.search(r"<element>(\d+)</element>", el)
        rows = parse_value(val)
        if rows is None:
            continue
        nr, nc = int(rows_decl.group(1)), int(cols_decl.group(1))
        truncated = not (len(rows) == nr and all(len(r) == nc for r in rows))
        items.append({"name": name, "line": int(line.group(1)) if line else -1,
                      "rows": nr, "cols": nc, "truncated": truncated, "value": rows})
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mlx_one_code.json")
    with open(dst, "w") as f:
        json.dump({"source": "utils/One_code.mlx: matlab/output.xml", "items": items}, f, indent=0)
    for it in items:
        print(it["name"], it["line"], it["rows"], it["cols"], "TRUNC" if it["truncated"] else "")


if __name__ == "__main__":
    main()
