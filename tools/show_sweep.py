#!/usr/bin/env python3
"""Print sweep .jsonl files (tools/sweep.py --out) as a table."""
import json
import sys

for f in sys.argv[1:]:
    print("==", f)
    for l in open(f):
        d = json.loads(l)
        if "error" in d:
            print(d)
            continue
        L = d["launch"]
        print(f'{d["point"]:52s} ms={d["ms"]:8.4f} frac={d["frac"]:.3f} gf={d["gflops"]:8.1f} same={int(d["same_as_first"])} '
              f'G={L["lanes_per_row"]} V={L["vec_elems"]} NT={L["reg_tiles"]} grid={L["grid"]} blk={L["block"]} smem={L["smem_bytes"]} '
              f'R={L["rows_per_slice"]} P={L["rows_per_warp"]} fl={L["reg_flavour"]} st={L["stages"]} cap={L["capacity"]} '
              f'passes={L["passes"]} mi={L["merge_items"]}')
