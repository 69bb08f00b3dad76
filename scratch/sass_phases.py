"""Static instruction budget of gn_step_kernel<2,float> per phase (no GPU needed):
    cuobjdump -xelf all dgpmp2_b200/lib/libdgpmp2_b200.so ; nvdisasm -gi c_abi.sm_100a.cubin > all_gi.txt
    python scratch/sass_phases.py all_gi.txt
Every SASS instruction is attributed to the call site in bcr_solve / gn_step_kernel found on its inlining chain
(nvdisasm --print-line-info-inline).  In the phases where one warp per scheduler has work the kernel issues in order
at ~4 cycles per dependent instruction, so the per-item instruction count IS the cost model."""
import collections
import re
import sys

KERNEL = '.text._ZN6dgpmp214gn_step_kernelILi2EfEE'
BCR_SITES = {631: 'elim, 1 lane/item', 632: 'elim, 4 lanes/item', 636: 'kept (Schur), 1 lane/item', 637: 'kept (Schur), 4 lanes/item',
             644: 'tail (block Thomas)', 655: 'back substitution, 1 lane/item', 656: 'back substitution, 4 lanes/item'}
KERNEL_SITES = {189: 'prologue (stage th)', 193: 'assembly (factors -> records)', 197: 'bcr_solve control',
                209: 'epilogue', 213: 'epilogue'}

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(KERNEL))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('.text.')), len(lines))
chain, counts, fp64, mem = [], collections.Counter(), collections.Counter(), collections.Counter()
fresh = True
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if not m:
        continue
    fresh = True
    op = m.group(1)
    phase = None
    for f, n in chain:
        if f == 'bcr.cuh' and n in BCR_SITES:
            phase = BCR_SITES[n]
    if phase is None:
        for f, n in chain:
            if f == 'kernels.cuh' and n in KERNEL_SITES:
                phase = KERNEL_SITES[n]
    phase = phase or 'other (kernel body)'
    counts[phase] += 1
    if op.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU')):
        fp64[phase] += 1
    if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'ATOMS', 'LD.', 'ST.')):
        mem[phase] += 1
print('%-34s %6s %6s %8s' % ('phase (static SASS, one copy of the code)', 'instr', 'fp64', 'ld/st'))
for k, v in sorted(counts.items(), key=lambda kv: -kv[1]):
    print('%-34s %6d %6d %8d' % (k, v, fp64[k], mem[k]))
print('%-34s %6d %6d %8d' % ('total', sum(counts.values()), sum(fp64.values()), sum(mem.values())))
