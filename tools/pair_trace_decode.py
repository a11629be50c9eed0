"""Decodes gpurun_out/pairtrace.txt (tools/pair_trace.py) into a per-iteration timeline."""
import sys
rows = [[int(x) for x in l.split()[1:]] for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pairtrace.txt")]
names = ['P0', "G4'", "G5'", 'P1', "G6'", "G7'", 'P2', 'G0', 'G1', 'P3', 'G2', 'G3']
for r in rows[8:11]:
    it, t = r[0], r[1:]
    base = t[12]
    print('it', it)
    print('  issued :', ' '.join('%s:%d' % (names[i], t[i] - base) for i in range(12)))
    print('  ready  :', ' '.join('%s:%d' % (names[i], t[12 + i] - base) for i in range(12)))
    print('  epi full v0..3:', [t[32 + v] - base for v in range(4)], 'handed:', [t[36 + v] - base for v in range(4)],
          'max back:', t[41] - base, 'h1: start', t[43] - base, 'computed', t[44] - base, 'fenced', t[45] - base,
          'arrived', t[42] - base)
print('iteration period:', [rows[i + 1][13] - rows[i][13] for i in range(5, 15)])
