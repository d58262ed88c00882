#!/usr/bin/env python
"""oracle/patch_reference.py <reference src dir> <out dir>: writes <out dir>/dwgsim.c = the reference's src/dwgsim.c with the
read-pair loop of dwgsim_core (`for (ii = 0; ii != n_pairs; ++ii, ++ctr) {...}`, src/dwgsim.c:636-1099) replaced by the call
into integration/dwgsim_b200_binding.c that INTEGRATION.md describes.  TEST INFRASTRUCTURE: the output lives under
oracle/_ref/ (git-ignored); no reference text is stored in this repository -- the loop is located by its header and removed
by brace matching."""
import re
import sys


def main():
    src_dir, out_dir = sys.argv[1], sys.argv[2]
    text = open(src_dir + "/dwgsim.c").read()
    m = re.search(r"for\s*\(\s*ii\s*=\s*0\s*;\s*ii\s*!=\s*n_pairs\s*;\s*\+\+ii\s*,\s*\+\+ctr\s*\)\s*\{", text)
    if not m:
        raise SystemExit("patch_reference: the core loop was not found")
    depth, i, in_str, in_chr, in_line, in_block = 1, m.end(), False, False, False, False
    while depth:
        c, nx = text[i], text[i + 1] if i + 1 < len(text) else ""
        if in_line:
            in_line = c != "\n"
        elif in_block:
            if c == "*" and nx == "/":
                in_block = False
                i += 1
        elif in_str:
            if c == "\\":
                i += 1
            elif c == '"':
                in_str = False
        elif in_chr:
            if c == "\\":
                i += 1
            elif c == "'":
                in_chr = False
        elif c == "/" and nx == "/":
            in_line = True
        elif c == "/" and nx == "*":
            in_block = True
        elif c == '"':
            in_str = True
        elif c == "'":
            in_chr = True
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
        i += 1
    call = ("(void)num_failed; (void)rand_ii; (void)ii;\n"
            "          dwgsim_b200_contig(opt, contig_i, name, &seq, mutseq[0], mutseq[1], n_pairs, regions_bed, l);\n"
            "          ctr += n_pairs; n_sim += n_pairs;")
    text = text[:m.start()] + call + text[i:]
    # the binding's header after the reference's own includes, its clean-up before the function's closing message
    text = text.replace('#include "dwgsim.h"', '#include "dwgsim.h"\n#include "dwgsim_b200_binding.h"', 1)
    done = re.search(r'fprintf\(stderr,\s*"\\n\[dwgsim_core\] Complete!\\n"\);', text)
    if not done:
        raise SystemExit("patch_reference: the end of dwgsim_core was not found")
    text = text[:done.start()] + "dwgsim_b200_close();\n  " + text[done.start():]
    open(out_dir + "/dwgsim.c", "w").write(text)


if __name__ == "__main__":
    main()
