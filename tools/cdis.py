#!/usr/bin/env python3
"""Throwaway JVM class-file disassembler for survey analysis (not part of the repo)."""
import struct, sys, zipfile
OPS = {}
def _init():
    names = """0 nop;1 aconst_null;2 iconst_m1;3 iconst_0;4 iconst_1;5 iconst_2;6 iconst_3;7 iconst_4;8 iconst_5;9 lconst_0;10 lconst_1;11 fconst_0;12 fconst_1;13 fconst_2;14 dconst_0;15 dconst_1;16 bipush b;17 sipush s;18 ldc c1;19 ldc_w c2;20 ldc2_w c2;21 iload b;22 lload b;23 fload b;24 dload b;25 aload b;26 iload_0;27 iload_1;28 iload_2;29 iload_3;30 lload_0;31 lload_1;32 lload_2;33 lload_3;34 fload_0;35 fload_1;36 fload_2;37 fload_3;38 dload_0;39 dload_1;40 dload_2;41 dload_3;42 aload_0;43 aload_1;44 aload_2;45 aload_3;46 iaload;47 laload;48 faload;49 daload;50 aaload;51 baload;52 caload;53 saload;54 istore b;55 lstore b;56 fstore b;57 dstore b;58 astore b;59 istore_0;60 istore_1;61 istore_2;62 istore_3;63 lstore_0;64 lstore_1;65 lstore_2;66 lstore_3;67 fstore_0;68 fstore_1;69 fstore_2;70 fstore_3;71 dstore_0;72 dstore_1;73 dstore_2;74 dstore_3;75 astore_0;76 astore_1;77 astore_2;78 astore_3;79 iastore;80 lastore;81 fastore;82 dastore;83 aastore;84 bastore;85 castore;86 sastore;87 pop;88 pop2;89 dup;90 dup_x1;91 dup_x2;92 dup2;93 dup2_x1;94 dup2_x2;95 swap;96 iadd;97 ladd;98 fadd;99 dadd;100 isub;101 lsub;102 fsub;103 dsub;104 imul;105 lmul;106 fmul;107 dmul;108 idiv;109 ldiv;110 fdiv;111 ddiv;112 irem;113 lrem;114 frem;115 drem;116 ineg;117 lneg;118 fneg;119 dneg;120 ishl;121 lshl;122 ishr;123 lshr;124 iushr;125 lushr;126 iand;127 land;128 ior;129 lor;130 ixor;131 lxor;132 iinc bb;133 i2l;134 i2f;135 i2d;136 l2i;137 l2f;138 l2d;139 f2i;140 f2l;141 f2d;142 d2i;143 d2l;144 d2f;145 i2b;146 i2c;147 i2s;148 lcmp;149 fcmpl;150 fcmpg;151 dcmpl;152 dcmpg;153 ifeq j;154 ifne j;155 iflt j;156 ifge j;157 ifgt j;158 ifle j;159 if_icmpeq j;160 if_icmpne j;161 if_icmplt j;162 if_icmpge j;163 if_icmpgt j;164 if_icmple j;165 if_acmpeq j;166 if_acmpne j;167 goto j;168 jsr j;169 ret b;170 tableswitch T;171 lookupswitch L;172 ireturn;173 lreturn;174 freturn;175 dreturn;176 areturn;177 return;178 getstatic c2;179 putstatic c2;180 getfield c2;181 putfield c2;182 invokevirtual c2;183 invokespecial c2;184 invokestatic c2;185 invokeinterface c2xx;186 invokedynamic c2xx;187 new c2;188 newarray b;189 anewarray c2;190 arraylength;191 athrow;192 checkcast c2;193 instanceof c2;194 monitorenter;195 monitorexit;196 wide W;197 multianewarray c2b;198 ifnull j;199 ifnonnull j;200 goto_w J;201 jsr_w J"""
    for ent in names.split(';'):
        p = ent.split()
        OPS[int(p[0])] = (p[1], p[2] if len(p) > 2 else '')
_init()
class CF:
    def __init__(s, data):
        s.d = data; s.p = 0
        assert s.u4() == 0xCAFEBABE
        s.u2(); s.u2()
        n = s.u2(); s.cp = [None]*n; i = 1
        while i < n:
            t = s.u1()
            if t == 1:
                l = s.u2(); s.cp[i] = ('utf', s.d[s.p:s.p+l].decode('utf-8', 'replace')); s.p += l
            elif t == 3: s.cp[i] = ('int', struct.unpack('>i', s.rd(4))[0])
            elif t == 4: s.cp[i] = ('float', struct.unpack('>f', s.rd(4))[0])
            elif t == 5: s.cp[i] = ('long', struct.unpack('>q', s.rd(8))[0]); i += 1
            elif t == 6: s.cp[i] = ('double', struct.unpack('>d', s.rd(8))[0]); i += 1
            elif t == 7: s.cp[i] = ('class', s.u2())
            elif t == 8: s.cp[i] = ('str', s.u2())
            elif t in (9, 10, 11): s.cp[i] = ('ref', s.u2(), s.u2())
            elif t == 12: s.cp[i] = ('nt', s.u2(), s.u2())
            elif t == 15: s.cp[i] = ('mh', s.u1(), s.u2())
            elif t == 16: s.cp[i] = ('mt', s.u2())
            elif t == 18: s.cp[i] = ('indy', s.u2(), s.u2())
            else: raise Exception('tag %d' % t)
            i += 1
        s.access = s.u2(); s.this = s.u2(); s.sup = s.u2()
        s.ifaces = [s.u2() for _ in range(s.u2())]
        s.fields = [s.member() for _ in range(s.u2())]
        s.methods = [s.member() for _ in range(s.u2())]
    def rd(s, n): r = s.d[s.p:s.p+n]; s.p += n; return r
    def u1(s): return s.rd(1)[0]
    def u2(s): return struct.unpack('>H', s.rd(2))[0]
    def u4(s): return struct.unpack('>I', s.rd(4))[0]
    def member(s):
        acc = s.u2(); name = s.cp[s.u2()][1]; desc = s.cp[s.u2()][1]; attrs = {}
        for _ in range(s.u2()):
            an = s.cp[s.u2()][1]; l = s.u4(); attrs[an] = s.rd(l)
        return (acc, name, desc, attrs)
    def c(s, i):
        e = s.cp[i]
        if e is None: return '?'
        t = e[0]
        if t == 'utf': return e[1]
        if t in ('int', 'float', 'long', 'double'): return '%s %r' % (t, e[1])
        if t == 'class': return s.cp[e[1]][1]
        if t == 'str': return 'String %r' % s.cp[e[1]][1]
        if t == 'ref': return '%s.%s' % (s.c(e[1]), s.c(e[2]))
        if t == 'nt': return '%s:%s' % (s.cp[e[1]][1], s.cp[e[2]][1])
        return repr(e)
    def dis(s, code):
        ms, ml, cl = struct.unpack('>HHI', code[:8]); b = code[8:8+cl]; pc = 0; out = []
        while pc < cl:
            op = b[pc]; name, fmt = OPS.get(op, ('op%d' % op, '')); start = pc; pc += 1; arg = ''
            if fmt == 'b': arg = str(b[pc]); pc += 1
            elif fmt == 's': arg = str(struct.unpack('>h', b[pc:pc+2])[0]); pc += 2
            elif fmt == 'c1': arg = s.c(b[pc]); pc += 1
            elif fmt == 'c2': arg = s.c(struct.unpack('>H', b[pc:pc+2])[0]); pc += 2
            elif fmt == 'c2xx': arg = s.c(struct.unpack('>H', b[pc:pc+2])[0]); pc += 4
            elif fmt == 'c2b': arg = s.c(struct.unpack('>H', b[pc:pc+2])[0]) + ' dim %d' % b[pc+2]; pc += 3
            elif fmt == 'bb': arg = '%d by %d' % (b[pc], struct.unpack('>b', b[pc+1:pc+2])[0]); pc += 2
            elif fmt == 'j': arg = '-> %d' % (start + struct.unpack('>h', b[pc:pc+2])[0]); pc += 2
            elif fmt == 'J': arg = '-> %d' % (start + struct.unpack('>i', b[pc:pc+4])[0]); pc += 4
            elif fmt == 'W':
                op2 = b[pc]; idx = struct.unpack('>H', b[pc+1:pc+3])[0]; pc += 3; arg = '%s %d' % (OPS[op2][0], idx)
                if op2 == 132: arg += ' by %d' % struct.unpack('>h', b[pc:pc+2])[0]; pc += 2
            elif fmt == 'T':
                pc = (pc + 3) & ~3; df, lo, hi = struct.unpack('>iii', b[pc:pc+12]); pc += 12
                offs = struct.unpack('>%di' % (hi-lo+1), b[pc:pc+4*(hi-lo+1)]); pc += 4*(hi-lo+1)
                arg = 'default->%d ' % (start+df) + ' '.join('%d->%d' % (lo+k, start+o) for k, o in enumerate(offs))
            elif fmt == 'L':
                pc = (pc + 3) & ~3; df, n = struct.unpack('>ii', b[pc:pc+8]); pc += 8; prs = []
                for _ in range(n):
                    k, o = struct.unpack('>ii', b[pc:pc+8]); pc += 8; prs.append('%d->%d' % (k, start+o))
                arg = 'default->%d ' % (start+df) + ' '.join(prs)
            out.append('%5d: %-16s %s' % (start, name, arg))
        return out
    def dump(s, only=None):
        print('class', s.c(s.this), 'extends', s.c(s.sup), 'implements', [s.c(i) for i in s.ifaces])
        for acc, n, d, a in s.fields:
            cv = ''
            if 'ConstantValue' in a: cv = ' = ' + s.c(struct.unpack('>H', a['ConstantValue'])[0])
            print('  field', n, d, cv)
        for acc, n, d, a in s.methods:
            if only and n not in only: continue
            print('  method %s%s acc=0x%x' % (n, d, acc))
            if 'Code' in a:
                for l in s.dis(a['Code']): print('   ', l)
if __name__ == '__main__':
    jar, cls = sys.argv[1], sys.argv[2]
    only = sys.argv[3].split(',') if len(sys.argv) > 3 else None
    z = zipfile.ZipFile(jar)
    CF(z.read(cls)).dump(only)
