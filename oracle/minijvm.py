"""minijvm — a small JVM bytecode interpreter, TEST INFRASTRUCTURE ONLY.

Why it exists: the reference's arithmetic lives in binary jars (`/root/reference/libs/lucene-core-5.3.1.jar`,
`fasta-utils-1.1.jar`) and this image has no JVM.  This module executes those class files directly —
unmodified bytes read from the jar with `zipfile` — so that golden vectors come from the reference's own
code (`tests/golden/make_lucene_golden.py` is the only caller; the vectors it writes are committed under
`tests/golden/lucene_*.json`).  Nothing in the product (`biospectra_b200/`) imports it.

Scope: the Java 7 instruction set without `jsr/ret/invokedynamic`, exact Java numerics (32/64-bit wrapping
integers, IEEE float rounding after every float operation, saturating f2i/d2i), exceptions with handler
tables, static initialisers, virtual/interface dispatch.  `java.*` is not available as bytecode (no rt.jar):
the few library classes Lucene touches on these paths are provided natively in `minijvm_rt.py`.

Value model: int-like → Python int (wrapped to 32 bits); long → Python int (64 bits) occupying two slots
(value, TOP) on the operand stack and in the locals exactly like the JVM, so that dup2/pop2 work without
type information; float → Python float already rounded to binary32; double → Python float (two slots);
references → None | str (java.lang.String) | JObject | JArray.
"""
from __future__ import annotations

import math
import struct
import sys
import zipfile

sys.setrecursionlimit(20000)

TOP = ("<top>",)  # filler slot of category-2 values

_pf = struct.Struct("<f")


def f32(x: float) -> float:
    """Round a Python float (binary64) to binary32, Java `(float)` semantics (overflow → ±Inf)."""
    try:
        return _pf.unpack(_pf.pack(x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


def i64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    return x - 0x10000000000000000 if x & 0x8000000000000000 else x


def float_to_raw_int_bits(f: float) -> int:
    return struct.unpack("<i", _pf.pack(f))[0]


def int_bits_to_float(i: int) -> float:
    return struct.unpack("<f", struct.pack("<i", i32(i)))[0]


def l2f(v: int) -> float:
    """long → float with a single rounding (float(v) would round to binary64 first)."""
    if v == 0:
        return 0.0
    s = -1.0 if v < 0 else 1.0
    a = abs(v)
    n = a.bit_length()
    if n <= 24:
        return s * float(a)
    sh = n - 24
    q, r = a >> sh, a & ((1 << sh) - 1)
    half = 1 << (sh - 1)
    if r > half or (r == half and (q & 1)):
        q += 1
    return s * math.ldexp(float(q), sh)


class JavaThrow(Exception):
    """A Java exception in flight; `.obj` is the Throwable instance."""

    def __init__(self, obj):
        self.obj = obj
        msg = obj.f.get("java/lang/Throwable.message") if isinstance(obj, JObject) else None
        super().__init__("%s: %s" % (obj.cls.name if isinstance(obj, JObject) else obj, msg))


class JObject:
    __slots__ = ("cls", "f", "p")

    def __init__(self, cls):
        self.cls = cls
        self.f = {}
        self.p = None  # payload of natively implemented classes

    def __repr__(self):
        return "<%s@%x>" % (self.cls.name, id(self) & 0xFFFFFF)


class JArray:
    __slots__ = ("etype", "a")

    def __init__(self, etype, a):
        self.etype = etype  # descriptor of the element type: 'I', 'F', 'Ljava/lang/String;', '[I' …
        self.a = a

    def __repr__(self):
        return "<%s[%d]>" % (self.etype, len(self.a))


class JMethod:
    __slots__ = ("cls", "name", "desc", "acc", "code", "max_locals", "exc", "nslots", "arg_cat", "ret", "fn", "is_static")

    def __init__(self, cls, name, desc, acc):
        self.cls, self.name, self.desc, self.acc = cls, name, desc, acc
        self.code = None
        self.max_locals = 0
        self.exc = ()
        self.fn = None  # native implementation: fn(vm, args) with args = [this?] + values (no TOP fillers)
        self.is_static = bool(acc & 0x0008)
        self.arg_cat, self.ret = parse_desc(desc)
        self.nslots = sum(self.arg_cat) + (0 if self.is_static else 1)

    def __repr__(self):
        return "%s.%s%s" % (self.cls.name, self.name, self.desc)


_DESC_CACHE = {}


def parse_desc(desc):
    """'(IJLfoo;[D)F' → ([1,2,1,1], 'F'): slot count per argument and the return descriptor."""
    r = _DESC_CACHE.get(desc)
    if r:
        return r
    i = 1
    cats = []
    while desc[i] != ")":
        c = desc[i]
        if c in "JD":
            cats.append(2)
            i += 1
        elif c == "L":
            cats.append(1)
            i = desc.index(";", i) + 1
        elif c == "[":
            while desc[i] == "[":
                i += 1
            if desc[i] == "L":
                i = desc.index(";", i) + 1
            else:
                i += 1
            cats.append(1)
        else:
            cats.append(1)
            i += 1
    r = (cats, desc[i + 1:])
    _DESC_CACHE[desc] = r
    return r


def default_of(desc):
    c = desc[0]
    if c in "IBCSZJ":
        return 0
    if c in "FD":
        return 0.0
    return None


class JClass:
    def __init__(self, vm, name):
        self.vm = vm
        self.name = name
        self.super_name = None
        self.sup = None
        self.iface_names = []
        self.acc = 0
        self.cp = []
        self.fields = {}  # name -> (acc, desc)
        self.methods = {}  # (name, desc) -> JMethod
        self.statics = {}
        self.init_state = 0  # 0 no, 1 running, 2 done
        self.native = False
        self._vcache = {}
        self._all_fields = None
        self._supers = None
        self.class_obj = None
        self.source = None

    def __repr__(self):
        return "JClass(%s)" % self.name

    # ---- hierarchy helpers
    def supers(self):
        """All classes and interfaces this class is assignable to (names)."""
        if self._supers is None:
            s = {self.name}
            if self.sup is not None:
                s |= self.sup.supers()
            for n in self.iface_names:
                s |= self.vm.load(n).supers()
            self._supers = s
        return self._supers

    def all_instance_fields(self):
        if self._all_fields is None:
            out = []
            c = self
            while c is not None:
                for n, (acc, d) in c.fields.items():
                    if not acc & 0x0008:
                        out.append((c.name + "." + n, default_of(d)))
                c = c.sup
            self._all_fields = out
        return self._all_fields

    def find_method(self, name, desc):
        """Virtual lookup from this class upwards; returns a JMethod with code or a native fn, else None."""
        key = (name, desc)
        m = self._vcache.get(key)
        if m is not None:
            return m
        c = self
        while c is not None:
            m = self.vm.native_method(c, name, desc)
            if m is not None:
                break
            m = c.methods.get(key)
            if m is not None and (m.code is not None or m.fn is not None):
                break
            m = None
            c = c.sup
        if m is None:
            # interfaces (natively implemented ones only: Java 7 interfaces have no code)
            for n in self.supers():
                ic = self.vm.load(n)
                m = self.vm.native_method(ic, name, desc)
                if m is not None:
                    break
        if m is not None:
            self._vcache[key] = m
        return m


class ClassFile:
    """Parser of the class-file format (JVMS §4) — constant pool, fields, methods with Code + exception table."""

    def __init__(self, data):
        self.d = data
        self.p = 0

    def rd(self, n):
        r = self.d[self.p:self.p + n]
        self.p += n
        return r

    def u1(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def u2(self):
        v = (self.d[self.p] << 8) | self.d[self.p + 1]
        self.p += 2
        return v

    def u4(self):
        return struct.unpack(">I", self.rd(4))[0]

    def parse_into(self, cls: JClass):
        assert self.u4() == 0xCAFEBABE
        self.u2()
        major = self.u2()
        n = self.u2()
        cp = [None] * n
        i = 1
        while i < n:
            t = self.u1()
            if t == 1:
                ln = self.u2()
                cp[i] = ("utf", _mutf8(self.rd(ln)))
            elif t == 3:
                cp[i] = ("int", struct.unpack(">i", self.rd(4))[0])
            elif t == 4:
                cp[i] = ("float", struct.unpack(">f", self.rd(4))[0])
            elif t == 5:
                cp[i] = ("long", struct.unpack(">q", self.rd(8))[0])
                i += 1
            elif t == 6:
                cp[i] = ("double", struct.unpack(">d", self.rd(8))[0])
                i += 1
            elif t == 7:
                cp[i] = ("class", self.u2())
            elif t == 8:
                cp[i] = ("str", self.u2())
            elif t in (9, 10, 11):
                cp[i] = ("ref", self.u2(), self.u2())
            elif t == 12:
                cp[i] = ("nt", self.u2(), self.u2())
            elif t == 15:
                cp[i] = ("mh", self.u1(), self.u2())
            elif t == 16:
                cp[i] = ("mt", self.u2())
            elif t == 18:
                cp[i] = ("indy", self.u2(), self.u2())
            else:
                raise ValueError("constant pool tag %d" % t)
            i += 1
        cls.cp = cp
        cls.major = major
        cls.acc = self.u2()
        this = cp[cp[self.u2()][1]][1]
        assert this == cls.name, (this, cls.name)
        s = self.u2()
        cls.super_name = cp[cp[s][1]][1] if s else None
        cls.iface_names = [cp[cp[self.u2()][1]][1] for _ in range(self.u2())]
        for _ in range(self.u2()):
            acc, name, desc, attrs = self.member(cp)
            cls.fields[name] = (acc, desc)
            if acc & 0x0008:
                v = default_of(desc)
                if "ConstantValue" in attrs:
                    e = cp[struct.unpack(">H", attrs["ConstantValue"])[0]]
                    v = cp[e[1]][1] if e[0] == "str" else e[1]
                cls.statics[name] = v
        for _ in range(self.u2()):
            acc, name, desc, attrs = self.member(cp)
            m = JMethod(cls, name, desc, acc)
            code = attrs.get("Code")
            if code is not None:
                ms, ml, cl = struct.unpack(">HHI", code[:8])
                m.code = code[8:8 + cl]
                m.max_locals = ml
                q = 8 + cl
                ne = struct.unpack(">H", code[q:q + 2])[0]
                q += 2
                exc = []
                for _e in range(ne):
                    s0, e0, h0, ct = struct.unpack(">HHHH", code[q:q + 8])
                    q += 8
                    exc.append((s0, e0, h0, cp[cp[ct][1]][1] if ct else None))
                m.exc = tuple(exc)
            cls.methods[(name, desc)] = m

    def member(self, cp):
        acc = self.u2()
        name = cp[self.u2()][1]
        desc = cp[self.u2()][1]
        attrs = {}
        for _ in range(self.u2()):
            an = cp[self.u2()][1]
            ln = self.u4()
            attrs[an] = self.rd(ln)
        return acc, name, desc, attrs


def _mutf8(b: bytes) -> str:
    try:
        return b.decode("utf-8")
    except UnicodeDecodeError:
        return b.decode("utf-8", "surrogatepass") if b"\xc0\x80" not in b else b.replace(b"\xc0\x80", b"\x00").decode("utf-8", "replace")


class VM:
    def __init__(self, jars):
        self.zips = [zipfile.ZipFile(j) for j in jars]
        self.classes = {}
        self.natives = {}  # (class, name, desc|None) -> fn
        self.native_supers = {}  # name -> (super, [ifaces]) for natively provided classes
        self.n_insn = 0
        self.trace = None
        self.executed = {}  # "Class.method" -> calls  (evidence of what really ran from the jar)
        import minijvm_rt
        minijvm_rt.install(self)

    # ---- class loading
    def load(self, name) -> JClass:
        c = self.classes.get(name)
        if c is not None:
            return c
        if name[0] == "[":
            c = JClass(self, name)
            c.super_name = "java/lang/Object"
            c.native = True
            c.iface_names = ["java/lang/Cloneable", "java/io/Serializable"]
            self.classes[name] = c
            c.sup = self.load("java/lang/Object")
            c.init_state = 2
            return c
        data = None
        for z in self.zips:
            try:
                data = z.read(name + ".class")
                break
            except KeyError:
                pass
        c = JClass(self, name)
        self.classes[name] = c
        if data is not None:
            ClassFile(data).parse_into(c)
            c.source = "jar"
        else:
            ns = self.native_supers.get(name)
            if ns is None:
                if not name.startswith(("java/", "javax/", "sun/")):
                    del self.classes[name]
                    raise KeyError("class not found in jars: " + name)
                ns = self.guess_native(name)
            c.native = True
            c.super_name, c.iface_names = ns[0], list(ns[1])
            c.source = "native"
            c.init_state = 2
        if c.super_name:
            c.sup = self.load(c.super_name)
        return c

    def guess_native(self, name):
        if name.endswith("Error"):
            return ("java/lang/Error", [])
        if name.endswith("Exception"):
            return ("java/lang/RuntimeException", [])
        return ("java/lang/Object", [])

    def define_native(self, name, sup="java/lang/Object", ifaces=()):
        self.native_supers[name] = (sup if name != "java/lang/Object" else None, list(ifaces))

    def native(self, cls, name, desc=None):
        def deco(fn):
            self.natives[(cls, name, desc)] = fn
            return fn
        return deco

    def native_method(self, c: JClass, name, desc):
        fn = self.natives.get((c.name, name, desc))
        if fn is None:
            fn = self.natives.get((c.name, name, None))
        if fn is None:
            return None
        nm = c.__dict__.setdefault("_native_methods", {})
        m = nm.get((name, desc))
        if m is None or m.fn is not fn:
            j = c.methods.get((name, desc))
            m = JMethod(c, name, desc, j.acc if j is not None else 0x0001)
            m.fn = fn
            nm[(name, desc)] = m
        return m

    def define_class(self, name, sup, ifaces=(), methods=None, fields=None):
        """A class implemented in Python that extends a jar class (stub readers / postings for the tests)."""
        c = JClass(self, name)
        c.super_name = sup
        c.iface_names = list(ifaces)
        c.source = "python"
        self.classes[name] = c
        c.sup = self.load(sup)
        c.init_state = 2
        for n, d in (fields or {}).items():
            c.fields[n] = (0, d)
        for (mn, md), fn in (methods or {}).items():
            m = JMethod(c, mn, md, 0x0001)
            m.fn = fn
            c.methods[(mn, md)] = m
        return c

    def init_class(self, c: JClass):
        if c.init_state:
            return
        c.init_state = 1
        if c.sup is not None:
            self.init_class(c.sup)
        fn = self.natives.get((c.name, "<clinit>", None))
        if fn is not None:
            fn(self, [])
        else:
            m = c.methods.get(("<clinit>", "()V"))
            if m is not None and m.code is not None:
                self.run(m, [])
        c.init_state = 2

    # ---- object helpers
    def new(self, cname) -> JObject:
        c = cname if isinstance(cname, JClass) else self.load(cname)
        self.init_class(c)
        o = JObject(c)
        f = o.f
        for k, v in c.all_instance_fields():
            f[k] = v
        return o

    def new_init(self, cname, desc, *args):
        o = self.new(cname)
        self.call_special(o.cls, "<init>", desc, o, *args)
        return o

    def class_of(self, ref) -> JClass:
        if isinstance(ref, JObject):
            return ref.cls
        if isinstance(ref, str):
            return self.load("java/lang/String")
        if isinstance(ref, JArray):
            return self.load("[" + ref.etype)
        raise TypeError("not a reference: %r" % (ref,))

    def instance_of(self, ref, cname) -> bool:
        if ref is None:
            return False
        if isinstance(ref, JArray):
            if cname in ("java/lang/Object", "java/lang/Cloneable", "java/io/Serializable"):
                return True
            if cname[0] != "[":
                return False
            et, ct = ref.etype, cname[1:]
            if et == ct:
                return True
            if et[0] == "L" and ct[0] == "L":
                return ct[1:-1] in self.load(et[1:-1]).supers()
            if et[0] == "[" and ct[0] == "[":
                return self.instance_of(JArray(et[1:], ()), ct)
            if ct == "Ljava/lang/Object;" and et[0] in "L[":
                return True
            return False
        return cname in self.class_of(ref).supers()

    def throw(self, cname, msg=None):
        o = self.new(cname)
        o.f["java/lang/Throwable.message"] = msg
        raise JavaThrow(o)

    def class_object(self, c: JClass) -> JObject:
        if c.class_obj is None:
            o = JObject(self.load("java/lang/Class"))
            o.p = c
            c.class_obj = o
        return c.class_obj

    # ---- calling conventions for Python callers (values, no TOP fillers)
    def _expand(self, m: JMethod, vals):
        out = []
        k = 0
        if not m.is_static:
            out.append(vals[0])
            k = 1
        for cat in m.arg_cat:
            out.append(vals[k])
            if cat == 2:
                out.append(TOP)
            k += 1
        assert k == len(vals), "argument count for %r" % m
        return out

    def invoke(self, m: JMethod, vals):
        """Invoke a resolved method with plain values; returns the plain result."""
        if m.fn is not None:
            return m.fn(self, list(vals))
        if m.code is None:
            raise RuntimeError("abstract or missing native: %r" % m)
        return self.run(m, self._expand(m, vals))

    def call_static(self, cname, name, desc, *args):
        c = self.load(cname)
        self.init_class(c)
        m = self.resolve_static(c, name, desc)
        return self.invoke(m, args)

    def call_virtual(self, obj, name, desc, *args):
        m = self.class_of(obj).find_method(name, desc)
        if m is None:
            raise RuntimeError("no method %s%s on %r" % (name, desc, obj))
        return self.invoke(m, (obj,) + args)

    def call_special(self, c: JClass, name, desc, obj, *args):
        m = self.resolve_special(c, name, desc)
        return self.invoke(m, (obj,) + args)

    def resolve_static(self, c, name, desc):
        k = c
        while k is not None:
            m = self.native_method(k, name, desc)
            if m is not None:
                return m
            m = k.methods.get((name, desc))
            if m is not None and (m.code is not None or m.fn is not None):
                return m
            k = k.sup
        raise RuntimeError("no static method %s.%s%s" % (c.name, name, desc))

    def resolve_special(self, c, name, desc):
        k = c
        while k is not None:
            m = self.native_method(k, name, desc)
            if m is not None:
                return m
            m = k.methods.get((name, desc))
            if m is not None and (m.code is not None or m.fn is not None):
                return m
            k = k.sup
        if name == "<init>" and c.native:
            return self._noop_init(c, desc)
        raise RuntimeError("no method %s.%s%s" % (c.name, name, desc))

    def _noop_init(self, c, desc):
        m = JMethod(c, "<init>", desc, 0x0001)

        def fn(vm, args):
            # generic constructor of natively stubbed classes: Throwable family keeps (message[, cause])
            o = args[0]
            if "java/lang/Throwable" in c.supers():
                for a in args[1:]:
                    if isinstance(a, str):
                        o.f["java/lang/Throwable.message"] = a
                    elif isinstance(a, JObject):
                        o.f["java/lang/Throwable.cause"] = a
            return None
        m.fn = fn
        c.methods[("<init>", desc)] = m
        return m

    # ---- field resolution
    def resolve_field(self, c: JClass, name):
        """→ (declaring JClass, is_static). Searches the class, its interfaces, then the superclass (JVMS §5.4.3.2)."""
        if c.native and c.init_state == 0:
            self.init_class(c)  # natively provided classes declare their static fields in their <clinit> stand-in
        k = c
        while k is not None:
            if name in k.fields:
                return k
            for n in k.iface_names:
                r = self._iface_field(self.load(n), name)
                if r is not None:
                    return r
            k = k.sup
        raise RuntimeError("no field %s.%s" % (c.name, name))

    def _iface_field(self, ic, name):
        if name in ic.fields:
            return ic
        for n in ic.iface_names:
            r = self._iface_field(self.load(n), name)
            if r is not None:
                return r
        return None

    # ---- the interpreter loop
    def run(self, m: JMethod, args):
        key = m.cls.name + "." + m.name
        ex = self.executed
        ex[key] = ex.get(key, 0) + 1
        code = m.code
        cls = m.cls
        cp = cls.cp
        loc = args + [None] * (m.max_locals - len(args))
        st = []
        push = st.append
        pop = st.pop
        pc = 0
        n_insn = 0
        while True:
            try:
                while True:
                    op = code[pc]
                    n_insn += 1
                    if op < 54:
                        if 26 <= op <= 45:  # xload_n
                            if op < 30:
                                push(loc[op - 26])
                            elif op < 34:
                                push(loc[op - 30])
                                push(TOP)
                            elif op < 38:
                                push(loc[op - 34])
                            elif op < 42:
                                push(loc[op - 38])
                                push(TOP)
                            else:
                                push(loc[op - 42])
                            pc += 1
                        elif 21 <= op <= 25:
                            push(loc[code[pc + 1]])
                            if op == 22 or op == 24:
                                push(TOP)
                            pc += 2
                        elif op >= 46:  # array loads
                            i = pop()
                            arr = pop()
                            if arr is None:
                                self.throw("java/lang/NullPointerException", "array load")
                            a = arr.a
                            if i < 0 or i >= len(a):
                                self.throw("java/lang/ArrayIndexOutOfBoundsException", str(i))
                            push(a[i])
                            if op == 47 or op == 49:
                                push(TOP)
                            pc += 1
                        elif op <= 8:
                            if op == 0:
                                pass
                            elif op == 1:
                                push(None)
                            else:
                                push(op - 3)
                            pc += 1
                        elif op <= 15:
                            if op <= 10:
                                push(op - 9)
                                push(TOP)
                            elif op <= 13:
                                push(float(op - 11))
                            else:
                                push(float(op - 14))
                                push(TOP)
                            pc += 1
                        elif op == 16:
                            v = code[pc + 1]
                            push(v - 256 if v > 127 else v)
                            pc += 2
                        elif op == 17:
                            v = (code[pc + 1] << 8) | code[pc + 2]
                            push(v - 65536 if v > 32767 else v)
                            pc += 3
                        elif op == 18 or op == 19:
                            if op == 18:
                                idx = code[pc + 1]
                                pc += 2
                            else:
                                idx = (code[pc + 1] << 8) | code[pc + 2]
                                pc += 3
                            e = cp[idx]
                            t = e[0]
                            if t == "int" or t == "float":
                                push(e[1])
                            elif t == "str":
                                push(cp[e[1]][1])
                            elif t == "class":
                                push(self.class_object(self.load(cp[e[1]][1])))
                            else:
                                raise RuntimeError("ldc of %r" % (e,))
                        elif op == 20:
                            idx = (code[pc + 1] << 8) | code[pc + 2]
                            push(cp[idx][1])
                            push(TOP)
                            pc += 3
                        else:
                            raise RuntimeError("opcode %d" % op)
                    elif op < 96:
                        if 59 <= op <= 78:  # xstore_n
                            if op < 63:
                                loc[op - 59] = pop()
                            elif op < 67:
                                pop()
                                loc[op - 63] = pop()
                                loc[op - 62] = TOP
                            elif op < 71:
                                loc[op - 67] = pop()
                            elif op < 75:
                                pop()
                                loc[op - 71] = pop()
                                loc[op - 70] = TOP
                            else:
                                loc[op - 75] = pop()
                            pc += 1
                        elif op <= 58:
                            n = code[pc + 1]
                            if op == 55 or op == 57:
                                pop()
                                loc[n] = pop()
                                loc[n + 1] = TOP
                            else:
                                loc[n] = pop()
                            pc += 2
                        elif op <= 86:  # array stores
                            if op == 80 or op == 82:
                                pop()
                            v = pop()
                            i = pop()
                            arr = pop()
                            if arr is None:
                                self.throw("java/lang/NullPointerException", "array store")
                            a = arr.a
                            if i < 0 or i >= len(a):
                                self.throw("java/lang/ArrayIndexOutOfBoundsException", str(i))
                            if op == 84:
                                if arr.etype == "Z":
                                    v &= 1
                                else:
                                    v &= 0xFF
                                    if v > 127:
                                        v -= 256
                            elif op == 85:
                                v &= 0xFFFF
                            elif op == 86:
                                v &= 0xFFFF
                                if v > 32767:
                                    v -= 65536
                            a[i] = v
                            pc += 1
                        elif op == 87:
                            pop()
                            pc += 1
                        elif op == 88:
                            pop()
                            pop()
                            pc += 1
                        elif op == 89:
                            push(st[-1])
                            pc += 1
                        elif op == 90:  # dup_x1
                            v1 = pop()
                            v2 = pop()
                            push(v1)
                            push(v2)
                            push(v1)
                            pc += 1
                        elif op == 91:  # dup_x2
                            v1 = pop()
                            v2 = pop()
                            v3 = pop()
                            push(v1)
                            push(v3)
                            push(v2)
                            push(v1)
                            pc += 1
                        elif op == 92:  # dup2
                            push(st[-2])
                            push(st[-2])
                            pc += 1
                        elif op == 93:  # dup2_x1
                            v1 = pop()
                            v2 = pop()
                            v3 = pop()
                            push(v2)
                            push(v1)
                            push(v3)
                            push(v2)
                            push(v1)
                            pc += 1
                        elif op == 94:  # dup2_x2
                            v1 = pop()
                            v2 = pop()
                            v3 = pop()
                            v4 = pop()
                            push(v2)
                            push(v1)
                            push(v4)
                            push(v3)
                            push(v2)
                            push(v1)
                            pc += 1
                        elif op == 95:
                            v1 = pop()
                            v2 = pop()
                            push(v1)
                            push(v2)
                            pc += 1
                        else:
                            raise RuntimeError("opcode %d" % op)
                    elif op < 153:
                        if op == 96:
                            b = pop()
                            push(i32(pop() + b))
                        elif op == 100:
                            b = pop()
                            push(i32(pop() - b))
                        elif op == 132:  # iinc
                            n = code[pc + 1]
                            d = code[pc + 2]
                            loc[n] = i32(loc[n] + (d - 256 if d > 127 else d))
                            pc += 3
                            continue
                        elif op == 98:
                            b = pop()
                            push(f32(pop() + b))
                        elif op == 106:
                            b = pop()
                            push(f32(pop() * b))
                        elif op == 102:
                            b = pop()
                            push(f32(pop() - b))
                        elif op == 110:
                            b = pop()
                            a = pop()
                            push(f32(_fdiv(a, b)))
                        elif op == 97 or op == 101 or op == 105:
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            push(i64(a + b if op == 97 else (a - b if op == 101 else a * b)))
                            push(TOP)
                        elif op == 99 or op == 103 or op == 107 or op == 111:
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            if op == 99:
                                r = a + b
                            elif op == 103:
                                r = a - b
                            elif op == 107:
                                r = _dmul(a, b)
                            else:
                                r = _fdiv(a, b)
                            push(r)
                            push(TOP)
                        elif op == 104:
                            b = pop()
                            push(i32(pop() * b))
                        elif op == 108 or op == 112:
                            b = pop()
                            a = pop()
                            if b == 0:
                                self.throw("java/lang/ArithmeticException", "/ by zero")
                            q = abs(a) // abs(b)
                            if (a < 0) != (b < 0):
                                q = -q
                            push(i32(q) if op == 108 else i32(a - q * b))
                        elif op == 109 or op == 113:
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            if b == 0:
                                self.throw("java/lang/ArithmeticException", "/ by zero")
                            q = abs(a) // abs(b)
                            if (a < 0) != (b < 0):
                                q = -q
                            push(i64(q) if op == 109 else i64(a - q * b))
                            push(TOP)
                        elif op == 114:
                            b = pop()
                            a = pop()
                            push(f32(_frem(a, b)))
                        elif op == 115:
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            push(_frem(a, b))
                            push(TOP)
                        elif op == 116:
                            push(i32(-pop()))
                        elif op == 117:
                            pop()
                            push(i64(-pop()))
                            push(TOP)
                        elif op == 118:
                            push(-pop())
                        elif op == 119:
                            pop()
                            push(-pop())
                            push(TOP)
                        elif op == 120:
                            b = pop()
                            push(i32(pop() << (b & 31)))
                        elif op == 122:
                            b = pop()
                            push(pop() >> (b & 31))
                        elif op == 124:
                            b = pop()
                            push(i32((pop() & 0xFFFFFFFF) >> (b & 31)))
                        elif op == 121 or op == 123 or op == 125:
                            b = pop()
                            pop()
                            a = pop()
                            b &= 63
                            if op == 121:
                                r = i64(a << b)
                            elif op == 123:
                                r = a >> b
                            else:
                                r = i64((a & 0xFFFFFFFFFFFFFFFF) >> b)
                            push(r)
                            push(TOP)
                        elif op == 126:
                            b = pop()
                            push(pop() & b)
                        elif op == 128:
                            b = pop()
                            push(pop() | b)
                        elif op == 130:
                            b = pop()
                            push(pop() ^ b)
                        elif op == 127 or op == 129 or op == 131:
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            push(a & b if op == 127 else (a | b if op == 129 else a ^ b))
                            push(TOP)
                        elif op == 133:  # i2l
                            push(TOP)
                        elif op == 134:  # i2f
                            push(f32(float(pop())))
                        elif op == 135:  # i2d
                            push(float(pop()))
                            push(TOP)
                        elif op == 136:  # l2i
                            pop()
                            push(i32(pop()))
                        elif op == 137:  # l2f
                            pop()
                            push(l2f(pop()))
                        elif op == 138:  # l2d
                            pop()
                            push(float(pop()))
                            push(TOP)
                        elif op == 139:  # f2i
                            push(_f2i(pop(), 31))
                        elif op == 140:  # f2l
                            push(_f2i(pop(), 63))
                            push(TOP)
                        elif op == 141:  # f2d
                            push(TOP)
                        elif op == 142:  # d2i
                            pop()
                            push(_f2i(pop(), 31))
                        elif op == 143:  # d2l
                            pop()
                            push(_f2i(pop(), 63))
                            push(TOP)
                        elif op == 144:  # d2f
                            pop()
                            push(f32(pop()))
                        elif op == 145:
                            v = pop() & 0xFF
                            push(v - 256 if v > 127 else v)
                        elif op == 146:
                            push(pop() & 0xFFFF)
                        elif op == 147:
                            v = pop() & 0xFFFF
                            push(v - 65536 if v > 32767 else v)
                        elif op == 148:  # lcmp
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            push((a > b) - (a < b))
                        elif op == 149 or op == 150:  # fcmpl / fcmpg
                            b = pop()
                            a = pop()
                            if a != a or b != b:
                                push(-1 if op == 149 else 1)
                            else:
                                push((a > b) - (a < b))
                        elif op == 151 or op == 152:
                            pop()
                            b = pop()
                            pop()
                            a = pop()
                            if a != a or b != b:
                                push(-1 if op == 151 else 1)
                            else:
                                push((a > b) - (a < b))
                        else:
                            raise RuntimeError("opcode %d" % op)
                        pc += 1
                    elif op < 172:
                        if op <= 158:
                            v = pop()
                            if op == 153:
                                t = v == 0
                            elif op == 154:
                                t = v != 0
                            elif op == 155:
                                t = v < 0
                            elif op == 156:
                                t = v >= 0
                            elif op == 157:
                                t = v > 0
                            else:
                                t = v <= 0
                        elif op <= 164:
                            b = pop()
                            a = pop()
                            if op == 159:
                                t = a == b
                            elif op == 160:
                                t = a != b
                            elif op == 161:
                                t = a < b
                            elif op == 162:
                                t = a >= b
                            elif op == 163:
                                t = a > b
                            else:
                                t = a <= b
                        elif op <= 166:
                            b = pop()
                            a = pop()
                            same = (a is b) or (isinstance(a, str) and isinstance(b, str) and a == b and _interned(a, b))
                            t = same if op == 165 else not same
                        elif op == 167:
                            t = True
                        elif op == 170:  # tableswitch
                            q = (pc + 4) & ~3
                            df, lo, hi = struct.unpack(">iii", code[q:q + 12])
                            v = pop()
                            if lo <= v <= hi:
                                off = struct.unpack(">i", code[q + 12 + 4 * (v - lo):q + 16 + 4 * (v - lo)])[0]
                            else:
                                off = df
                            pc += off
                            continue
                        elif op == 171:  # lookupswitch
                            q = (pc + 4) & ~3
                            df, np_ = struct.unpack(">ii", code[q:q + 8])
                            v = pop()
                            off = df
                            for k in range(np_):
                                mk, mo = struct.unpack(">ii", code[q + 8 + 8 * k:q + 16 + 8 * k])
                                if mk == v:
                                    off = mo
                                    break
                            pc += off
                            continue
                        else:
                            raise RuntimeError("opcode %d (jsr/ret unsupported)" % op)
                        if t:
                            off = (code[pc + 1] << 8) | code[pc + 2]
                            pc += off - 65536 if off > 32767 else off
                        else:
                            pc += 3
                    elif op <= 177:  # returns
                        self.n_insn += n_insn
                        if op == 177:
                            return None
                        if op == 173 or op == 175:
                            pop()
                        return pop()
                    elif op <= 181:  # field access
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        e = cp[idx]
                        if len(e) == 3:  # unresolved ('ref', class, nt)
                            cname = cp[cp[e[1]][1]][1]
                            nt = cp[e[2]]
                            fname, fdesc = cp[nt[1]][1], cp[nt[2]][1]
                            decl = self.resolve_field(self.load(cname), fname)
                            e = ("ref", decl, fname, decl.name + "." + fname, 2 if fdesc[0] in "JD" else 1)
                            cp[idx] = e
                        if op == 178:
                            decl = e[1]
                            if decl.init_state == 0:
                                self.init_class(decl)
                            push(decl.statics[e[2]])
                            if e[4] == 2:
                                push(TOP)
                        elif op == 179:
                            decl = e[1]
                            if decl.init_state == 0:
                                self.init_class(decl)
                            if e[4] == 2:
                                pop()
                            decl.statics[e[2]] = pop()
                        elif op == 180:
                            o = pop()
                            if o is None:
                                self.throw("java/lang/NullPointerException", "getfield " + e[3])
                            push(o.f[e[3]])
                            if e[4] == 2:
                                push(TOP)
                        else:
                            if e[4] == 2:
                                pop()
                            v = pop()
                            o = pop()
                            if o is None:
                                self.throw("java/lang/NullPointerException", "putfield " + e[3])
                            o.f[e[3]] = v
                        pc += 3
                    elif op <= 185:  # invokes
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        e = cp[idx]
                        if len(e) == 3:
                            cname = cp[cp[e[1]][1]][1]
                            nt = cp[e[2]]
                            mname, mdesc = cp[nt[1]][1], cp[nt[2]][1]
                            cats, ret = parse_desc(mdesc)
                            e = ("ref", cname, mname, mdesc, sum(cats), None, 0 if ret == "V" else (2 if ret in "JD" else 1))
                            cp[idx] = e
                        nsl = e[4]
                        if op == 184:
                            tm = e[5]
                            if tm is None:
                                c = self.load(e[1])
                                self.init_class(c)
                                tm = self.resolve_static(c, e[2], e[3])
                                cp[idx] = e[:5] + (tm,) + e[6:]
                            if nsl:
                                a = st[-nsl:]
                                del st[-nsl:]
                            else:
                                a = []
                        else:
                            nsl += 1
                            a = st[-nsl:]
                            del st[-nsl:]
                            o = a[0]
                            if o is None:
                                self.throw("java/lang/NullPointerException", "invoke %s.%s" % (e[1], e[2]))
                            if op == 183:
                                tm = e[5]
                                if tm is None:
                                    tm = self.resolve_special(self.load(e[1]), e[2], e[3])
                                    cp[idx] = e[:5] + (tm,) + e[6:]
                            else:
                                oc = o.cls if o.__class__ is JObject else self.class_of(o)
                                tm = oc._vcache.get((e[2], e[3])) or oc.find_method(e[2], e[3])
                                if tm is None:
                                    raise RuntimeError("no method %s.%s%s on %r (called from %r)" % (e[1], e[2], e[3], o, m))
                        if tm.fn is not None:
                            if TOP in a:
                                a = [x for x in a if x is not TOP]
                            r = tm.fn(self, a)
                        elif tm.code is not None:
                            r = self.run(tm, a)
                        else:
                            raise RuntimeError("abstract method %r invoked from %r" % (tm, m))
                        rc = e[6]
                        if rc:
                            push(r)
                            if rc == 2:
                                push(TOP)
                        pc += 5 if op == 185 else 3
                    elif op == 187:
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        push(self.new(cp[cp[idx][1]][1]))
                        pc += 3
                    elif op == 188:
                        t = code[pc + 1]
                        n = pop()
                        if n < 0:
                            self.throw("java/lang/NegativeArraySizeException", str(n))
                        et = {4: "Z", 5: "C", 6: "F", 7: "D", 8: "B", 9: "S", 10: "I", 11: "J"}[t]
                        push(JArray(et, [0.0 if et in "FD" else 0] * n))
                        pc += 2
                    elif op == 189:
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        cn = cp[cp[idx][1]][1]
                        n = pop()
                        if n < 0:
                            self.throw("java/lang/NegativeArraySizeException", str(n))
                        push(JArray(cn if cn[0] == "[" else "L" + cn + ";", [None] * n))
                        pc += 3
                    elif op == 190:
                        arr = pop()
                        if arr is None:
                            self.throw("java/lang/NullPointerException", "arraylength")
                        push(len(arr.a))
                        pc += 1
                    elif op == 191:
                        o = pop()
                        if o is None:
                            self.throw("java/lang/NullPointerException", "athrow")
                        raise JavaThrow(o)
                    elif op == 192:
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        o = st[-1]
                        if o is not None:
                            cn = cp[cp[idx][1]][1]
                            if not self.instance_of(o, cn):
                                self.throw("java/lang/ClassCastException", "%r to %s" % (o, cn))
                        pc += 3
                    elif op == 193:
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        push(1 if self.instance_of(pop(), cp[cp[idx][1]][1]) else 0)
                        pc += 3
                    elif op == 194 or op == 195:
                        pop()
                        pc += 1
                    elif op == 196:  # wide
                        op2 = code[pc + 1]
                        n = (code[pc + 2] << 8) | code[pc + 3]
                        if op2 == 132:
                            d = (code[pc + 4] << 8) | code[pc + 5]
                            loc[n] = i32(loc[n] + (d - 65536 if d > 32767 else d))
                            pc += 6
                        elif 21 <= op2 <= 25:
                            push(loc[n])
                            if op2 == 22 or op2 == 24:
                                push(TOP)
                            pc += 4
                        elif 54 <= op2 <= 58:
                            if op2 == 55 or op2 == 57:
                                pop()
                                loc[n] = pop()
                                loc[n + 1] = TOP
                            else:
                                loc[n] = pop()
                            pc += 4
                        else:
                            raise RuntimeError("wide %d" % op2)
                    elif op == 197:  # multianewarray
                        idx = (code[pc + 1] << 8) | code[pc + 2]
                        dims = code[pc + 3]
                        cn = cp[cp[idx][1]][1]
                        counts = [pop() for _ in range(dims)][::-1]
                        push(_multi(cn, counts))
                        pc += 4
                    elif op == 198 or op == 199:
                        v = pop()
                        if (v is None) == (op == 198):
                            off = (code[pc + 1] << 8) | code[pc + 2]
                            pc += off - 65536 if off > 32767 else off
                        else:
                            pc += 3
                    elif op == 200:
                        pc += struct.unpack(">i", code[pc + 1:pc + 5])[0]
                    else:
                        raise RuntimeError("opcode %d" % op)
            except (RuntimeError, KeyError, TypeError, AttributeError, IndexError, AssertionError) as e:
                if hasattr(e, "add_note"):
                    e.add_note("  at %r pc=%d" % (m, pc))
                raise
            except JavaThrow as t:
                h = None
                for s0, e0, h0, ct in m.exc:
                    if s0 <= pc < e0 and (ct is None or self.instance_of(t.obj, ct)):
                        h = h0
                        break
                if h is None:
                    self.n_insn += n_insn
                    n_insn = 0
                    raise
                del st[:]
                push(t.obj)
                pc = h


_INTERN = {}


def _interned(a, b):
    # String identity: constants are interned; Python strs stand for Java Strings, so `==` on equal strings is
    # taken as identity (Lucene compares field names with `==` after interning them).
    return True


def _fdiv(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def _dmul(a, b):
    try:
        return a * b
    except OverflowError:
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def _frem(a, b):
    if a != a or b != b or math.isinf(a) or b == 0.0:
        return math.nan
    if math.isinf(b):
        return a
    return math.fmod(a, b)


def _f2i(v, bits):
    if v != v:
        return 0
    lo, hi = -(1 << bits), (1 << bits) - 1
    if v >= hi:
        return hi
    if v <= lo:
        return lo
    return int(v)


def _multi(cn, counts):
    et = cn[1:]
    n = counts[0]
    if len(counts) == 1:
        if et[0] in "L[":
            return JArray(et, [None] * n)
        return JArray(et, [0.0 if et in "FD" else 0] * n)
    return JArray(et, [_multi(et, counts[1:]) for _ in range(n)])
