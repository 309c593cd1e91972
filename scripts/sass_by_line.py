"""Attribute the SASS instructions of one kernel to generated source lines / statement kinds.
usage: sass_by_line.py all_lines.sass(nvdisasm -g -c of the cubin) dumped_source.cu kernel_name"""
import re, collections, sys
sass, srcf, kern = sys.argv[1:4]
cnt = collections.Counter(); cur = None; on = False
for l in open(sass):
    if l.startswith(".text."):
        on = l.strip() == f".text.{kern}:"
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = int(m.group(2)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        cnt[cur] += 1
src = open(srcf).read().split('\n')
tot = sum(cnt.values()); print('total', tot, 'source lines', len(cnt))
cat = collections.Counter(); nl = collections.Counter()
for ln, c in cnt.items():
    t = src[ln - 1].strip() if ln and ln <= len(src) else '?'
    if 'VA_RCP' in t: k = 'rcp'
    elif 'VA_SQRT' in t: k = 'sqrt'
    elif re.search(r'\bexp\(', t): k = 'exp'
    elif re.search(r'\blog\(', t): k = 'log'
    elif re.search(r'\bpow\(', t): k = 'pow'
    elif '?' in t: k = 'select'
    elif t.startswith('if') or t.startswith('const int c') or t.startswith('} else'): k = 'cond'
    elif 'CACHE_LD' in t: k = 'cacheld'
    elif 'VA_CHUNK' in t: k = 'chunk'
    elif 'OUT_' in t: k = 'out'
    elif 'acc' in t: k = 'acc'
    elif re.search(r'__d\d', t): k = 'arith_deriv'
    else: k = 'arith_value'
    cat[k] += c; nl[k] += 1
for k, c in cat.most_common(): print(f'{k:12s} {c:6d} {100*c/tot:5.1f}%  over {nl[k]} lines')
for ln, c in cnt.most_common(15): print(c, ln, (src[ln - 1].strip()[:150] if ln else None))
