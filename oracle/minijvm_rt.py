"""Native stand-ins for the handful of `java.*` classes the executed Lucene / fasta-utils paths touch.
TEST INFRASTRUCTURE ONLY (see minijvm.py).  There is no rt.jar in this image, so these are written from the
Java SE API contracts; collections are kept tiny and simple (linear search with Java `equals` semantics,
insertion-ordered iteration) because the executed paths only hold a few elements in them.  Everything numeric
(`Math`, `Float`, `Double`, `Integer`, `Long`) follows the Java language/lib specification bit for bit:
`Math.sqrt` is IEEE correctly rounded; `Math.log` is glibc's (≤ 1 ulp, as the Java spec allows) — the one place
where a real JVM may differ in the last bit of a double, invisible after Lucene's `(float)` casts except on a
rounding boundary (the golden generator records every `Math.log` result it served so this can be audited).
"""
from __future__ import annotations

import math
import struct
import time

from minijvm import (JArray, JObject, JavaThrow, f32, float_to_raw_int_bits, i32, i64, int_bits_to_float)

OBJ = "java/lang/Object"


def jstr(vm, o) -> str:
    """String.valueOf(Object)."""
    if o is None:
        return "null"
    if isinstance(o, str):
        return o
    if isinstance(o, JObject):
        return vm.call_virtual(o, "toString", "()Ljava/lang/String;")
    return repr(o)


def jequals(vm, a, b) -> bool:
    if a is b:
        return True
    if a is None or b is None:
        return False
    if isinstance(a, str):
        return isinstance(b, str) and a == b
    if isinstance(a, JObject):
        return bool(vm.call_virtual(a, "equals", "(Ljava/lang/Object;)Z", b))
    return False


def jhash(vm, a) -> int:
    if a is None:
        return 0
    if isinstance(a, str):
        return string_hash(a)
    if isinstance(a, JObject):
        return vm.call_virtual(a, "hashCode", "()I")
    return id(a) & 0x7FFFFFFF


def string_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return i32(h)


def box(vm, cname, v):
    o = JObject(vm.load(cname))
    o.p = v
    return o


def mk(vm, cname, payload):
    o = JObject(vm.load(cname))
    o.p = payload
    return o


def make_iter(vm, seq):
    return mk(vm, "java/util/PyIterator", [list(seq), 0, None])


def java_double_to_string(d: float) -> str:
    if d != d:
        return "NaN"
    if math.isinf(d):
        return "Infinity" if d > 0 else "-Infinity"
    r = repr(d)
    if "e" in r or "E" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return "%sE%d" % (m, int(e))
    if "." not in r:
        r += ".0"
    return r


def java_float_to_string(f: float) -> str:
    import numpy as np
    if f != f or math.isinf(f):
        return java_double_to_string(f)
    r = str(np.float32(f))
    if "e" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return "%sE%d" % (m, int(e))
    if "." not in r:
        r += ".0"
    return r


def install(vm):
    N = vm.native
    D = vm.define_native
    vm.math_log_calls = []

    # ------------------------------------------------------------------ class skeletons
    D(OBJ, None)
    for n in ("java/lang/Cloneable", "java/io/Serializable", "java/lang/Comparable", "java/lang/CharSequence",
              "java/lang/Iterable", "java/util/Collection", "java/util/List", "java/util/Set", "java/util/Map",
              "java/util/Iterator", "java/util/Comparator", "java/util/Map$Entry", "java/io/Closeable",
              "java/lang/AutoCloseable", "java/lang/Runnable", "java/util/RandomAccess", "java/util/Queue",
              "java/util/SortedMap", "java/util/SortedSet", "java/util/concurrent/Callable", "java/lang/Appendable",
              "java/util/concurrent/locks/Lock", "java/util/concurrent/Executor", "java/util/concurrent/ExecutorService"):
        D(n)
    D("java/util/Collection", OBJ, ["java/lang/Iterable"])
    D("java/util/List", OBJ, ["java/util/Collection"])
    D("java/util/Set", OBJ, ["java/util/Collection"])
    D("java/lang/String", OBJ, ["java/lang/Comparable", "java/lang/CharSequence", "java/io/Serializable"])
    D("java/lang/Number", OBJ, ["java/io/Serializable"])
    for n in ("Integer", "Long", "Float", "Double", "Short", "Byte"):
        D("java/lang/" + n, "java/lang/Number", ["java/lang/Comparable"])
    D("java/lang/Boolean", OBJ, ["java/lang/Comparable"])
    D("java/lang/Character", OBJ, ["java/lang/Comparable"])
    D("java/lang/Throwable", OBJ, ["java/io/Serializable"])
    D("java/lang/Exception", "java/lang/Throwable")
    D("java/lang/Error", "java/lang/Throwable")
    D("java/lang/RuntimeException", "java/lang/Exception")
    D("java/io/IOException", "java/lang/Exception")
    D("java/io/EOFException", "java/io/IOException")
    D("java/io/FileNotFoundException", "java/io/IOException")
    D("java/lang/AssertionError", "java/lang/Error")
    D("java/lang/IndexOutOfBoundsException", "java/lang/RuntimeException")
    D("java/lang/ArrayIndexOutOfBoundsException", "java/lang/IndexOutOfBoundsException")
    D("java/lang/StringIndexOutOfBoundsException", "java/lang/IndexOutOfBoundsException")
    D("java/lang/IllegalArgumentException", "java/lang/RuntimeException")
    D("java/lang/NumberFormatException", "java/lang/IllegalArgumentException")
    D("java/util/AbstractCollection", OBJ, ["java/util/Collection"])
    D("java/util/AbstractList", "java/util/AbstractCollection", ["java/util/List"])
    D("java/util/AbstractSet", "java/util/AbstractCollection", ["java/util/Set"])
    D("java/util/AbstractMap", OBJ, ["java/util/Map"])
    D("java/util/ArrayList", "java/util/AbstractList", ["java/util/List", "java/util/RandomAccess"])
    D("java/util/LinkedList", "java/util/ArrayList", ["java/util/Queue"])
    D("java/util/Arrays$ArrayList", "java/util/ArrayList")
    D("java/util/HashMap", "java/util/AbstractMap", ["java/util/Map"])
    D("java/util/LinkedHashMap", "java/util/HashMap")
    D("java/util/IdentityHashMap", "java/util/HashMap")
    D("java/util/WeakHashMap", "java/util/HashMap")
    D("java/util/TreeMap", "java/util/HashMap", ["java/util/SortedMap"])
    D("java/util/concurrent/ConcurrentHashMap", "java/util/HashMap")
    D("java/util/HashSet", "java/util/AbstractSet", ["java/util/Set"])
    D("java/util/LinkedHashSet", "java/util/HashSet")
    D("java/util/TreeSet", "java/util/HashSet", ["java/util/SortedSet"])
    D("java/util/PyIterator", OBJ, ["java/util/Iterator"])
    D("java/util/PyEntry", OBJ, ["java/util/Map$Entry"])
    D("java/lang/StringBuilder", OBJ, ["java/lang/CharSequence", "java/lang/Appendable"])
    D("java/lang/StringBuffer", "java/lang/StringBuilder")
    D("java/io/Reader", OBJ, ["java/io/Closeable"])
    D("java/io/InputStream", OBJ, ["java/io/Closeable"])
    D("java/io/FileInputStream", "java/io/InputStream")
    D("java/io/InputStreamReader", "java/io/Reader")
    D("java/io/FileReader", "java/io/InputStreamReader")
    D("java/io/BufferedReader", "java/io/Reader")
    D("java/io/StringReader", "java/io/Reader")

    # ------------------------------------------------------------------ Object / Class / System
    @N(OBJ, "<init>")
    def _(vm, a):
        return None

    @N(OBJ, "hashCode")
    def _(vm, a):
        return id(a[0]) & 0x7FFFFFFF

    @N(OBJ, "equals")
    def _(vm, a):
        return 1 if a[0] is a[1] else 0

    @N(OBJ, "getClass")
    def _(vm, a):
        return vm.class_object(vm.class_of(a[0]))

    @N(OBJ, "toString")
    def _(vm, a):
        return "%s@%x" % (vm.class_of(a[0]).name.replace("/", "."), id(a[0]) & 0xFFFFFF)

    @N(OBJ, "clone")
    def _(vm, a):
        o = a[0]
        if isinstance(o, JArray):
            return JArray(o.etype, list(o.a))
        c = JObject(o.cls)
        c.f = dict(o.f)
        c.p = o.p
        return c

    @N("java/lang/Class", "desiredAssertionStatus")
    def _(vm, a):
        return 1 if getattr(vm, "assertions", False) else 0

    @N("java/lang/Class", "getName")
    def _(vm, a):
        return a[0].p.name.replace("/", ".")

    @N("java/lang/Class", "getSimpleName")
    def _(vm, a):
        return a[0].p.name.split("/")[-1].split("$")[-1]

    @N("java/lang/Class", "isInstance")
    def _(vm, a):
        return 1 if vm.instance_of(a[1], a[0].p.name) else 0

    @N("java/lang/Class", "isAssignableFrom")
    def _(vm, a):
        return 1 if a[0].p.name in a[1].p.supers() else 0

    @N("java/lang/Class", "hashCode")
    def _(vm, a):
        return string_hash(a[0].p.name)

    @N("java/lang/Class", "getClassLoader")
    def _(vm, a):
        return None

    @N("java/lang/Class", "getComponentType")
    def _(vm, a):
        n = a[0].p.name
        if n[0] != "[":
            return None
        et = n[1:]
        return vm.class_object(vm.load(et[1:-1] if et[0] == "L" else et))

    @N("java/lang/System", "arraycopy")
    def _(vm, a):
        src, sp, dst, dp, n = a
        if src is None or dst is None:
            vm.throw("java/lang/NullPointerException", "arraycopy")
        if n < 0 or sp < 0 or dp < 0 or sp + n > len(src.a) or dp + n > len(dst.a):
            vm.throw("java/lang/ArrayIndexOutOfBoundsException", "arraycopy")
        dst.a[dp:dp + n] = src.a[sp:sp + n]

    @N("java/lang/System", "currentTimeMillis")
    def _(vm, a):
        return int(time.time() * 1000)

    @N("java/lang/System", "nanoTime")
    def _(vm, a):
        return time.monotonic_ns()

    @N("java/lang/System", "getProperty")
    def _(vm, a):
        props = {"java.version": "1.7.0_80", "java.vm.name": "minijvm", "os.name": "Linux", "os.arch": "amd64",
                 "java.vendor": "minijvm", "sun.arch.data.model": "64", "line.separator": "\n", "file.separator": "/"}
        v = props.get(a[0])
        return v if v is not None or len(a) < 2 else a[1]

    @N("java/lang/System", "identityHashCode")
    def _(vm, a):
        return 0 if a[0] is None else id(a[0]) & 0x7FFFFFFF

    @N("java/lang/System", "<clinit>")
    def _(vm, a):
        c = vm.load("java/lang/System")
        c.fields["out"] = (0x0009, "Ljava/io/PrintStream;")
        c.fields["err"] = (0x0009, "Ljava/io/PrintStream;")
        c.statics["out"] = mk(vm, "java/io/PrintStream", None)
        c.statics["err"] = mk(vm, "java/io/PrintStream", None)
    vm.load("java/lang/System").init_state = 0

    @N("java/io/PrintStream", "println")
    def _(vm, a):
        print("[java]", *[jstr(vm, x) if not isinstance(x, (int, float)) else x for x in a[1:]])

    @N("java/io/PrintStream", "print")
    def _(vm, a):
        print("[java]", *[jstr(vm, x) if not isinstance(x, (int, float)) else x for x in a[1:]], end="")

    # ------------------------------------------------------------------ Throwable
    D("java/lang/Throwable", OBJ)
    tcls = vm.load("java/lang/Throwable")
    tcls.fields["message"] = (0, "Ljava/lang/String;")
    tcls.fields["cause"] = (0, "Ljava/lang/Throwable;")

    @N("java/lang/Throwable", "getMessage")
    def _(vm, a):
        return a[0].f.get("java/lang/Throwable.message")

    @N("java/lang/Throwable", "getCause")
    def _(vm, a):
        return a[0].f.get("java/lang/Throwable.cause")

    @N("java/lang/Throwable", "toString")
    def _(vm, a):
        m = a[0].f.get("java/lang/Throwable.message")
        n = a[0].cls.name.replace("/", ".")
        return n if m is None else n + ": " + m

    @N("java/lang/Throwable", "addSuppressed")
    def _(vm, a):
        return None

    @N("java/lang/Throwable", "initCause")
    def _(vm, a):
        a[0].f["java/lang/Throwable.cause"] = a[1]
        return a[0]

    @N("java/lang/Throwable", "fillInStackTrace")
    def _(vm, a):
        return a[0]

    @N("java/lang/Throwable", "printStackTrace")
    def _(vm, a):
        print("[java] exception:", jstr(vm, a[0]))

    # ------------------------------------------------------------------ Math
    for cn in ("java/lang/Math", "java/lang/StrictMath"):
        @N(cn, "sqrt")
        def _(vm, a):
            return math.sqrt(a[0]) if a[0] >= 0 else math.nan

        @N(cn, "log")
        def _(vm, a):
            x = a[0]
            if x != x or x < 0:
                r = math.nan
            elif x == 0:
                r = -math.inf
            elif math.isinf(x):
                r = math.inf
            else:
                r = math.log(x)
            vm.math_log_calls.append((x, r))
            return r

        @N(cn, "min")
        def _(vm, a):
            x, y = a
            if isinstance(x, float):
                if x != x or y != y:
                    return math.nan
                if x == 0.0 and y == 0.0:
                    return x if math.copysign(1, x) < 0 else y
            return x if x <= y else y

        @N(cn, "max")
        def _(vm, a):
            x, y = a
            if isinstance(x, float):
                if x != x or y != y:
                    return math.nan
                if x == 0.0 and y == 0.0:
                    return x if math.copysign(1, x) > 0 else y
            return x if x >= y else y

        @N(cn, "abs")
        def _(vm, a):
            x = a[0]
            if isinstance(x, float):
                return math.fabs(x)
            return x if x >= 0 else -x  # caller wraps; Integer.MIN_VALUE stays (wrap below)

        @N(cn, "floor")
        def _(vm, a):
            x = a[0]
            return x if x != x or math.isinf(x) else float(math.floor(x))

        @N(cn, "ceil")
        def _(vm, a):
            x = a[0]
            return x if x != x or math.isinf(x) else float(math.ceil(x))

        @N(cn, "pow")
        def _(vm, a):
            try:
                return math.pow(a[0], a[1])
            except (OverflowError, ValueError):
                return math.inf

        @N(cn, "exp")
        def _(vm, a):
            try:
                return math.exp(a[0])
            except OverflowError:
                return math.inf

        @N(cn, "round", "(F)I")
        def _(vm, a):
            x = a[0]
            if x != x:
                return 0
            return max(-(1 << 31), min((1 << 31) - 1, math.floor(x + 0.5)))

        @N(cn, "round", "(D)J")
        def _(vm, a):
            x = a[0]
            if x != x:
                return 0
            if math.isinf(x):
                return (1 << 63) - 1 if x > 0 else -(1 << 63)
            return max(-(1 << 63), min((1 << 63) - 1, math.floor(x + 0.5)))

    # ------------------------------------------------------------------ Float / Double / Integer / Long / Boolean / Character
    @N("java/lang/Float", "floatToRawIntBits")
    def _(vm, a):
        return float_to_raw_int_bits(a[0])

    @N("java/lang/Float", "floatToIntBits")
    def _(vm, a):
        return 0x7FC00000 if a[0] != a[0] else float_to_raw_int_bits(a[0])

    @N("java/lang/Float", "intBitsToFloat")
    def _(vm, a):
        return int_bits_to_float(a[0])

    @N("java/lang/Float", "isNaN")
    def _(vm, a):
        v = a[-1] if not isinstance(a[-1], JObject) else a[-1].p
        return 1 if v != v else 0

    @N("java/lang/Float", "isInfinite")
    def _(vm, a):
        v = a[-1] if not isinstance(a[-1], JObject) else a[-1].p
        return 1 if math.isinf(v) else 0

    @N("java/lang/Float", "compare")
    def _(vm, a):
        return _fcompare(a[0], a[1])

    @N("java/lang/Float", "valueOf", "(F)Ljava/lang/Float;")
    def _(vm, a):
        return box(vm, "java/lang/Float", a[0])

    @N("java/lang/Float", "<init>", "(F)V")
    def _(vm, a):
        a[0].p = a[1]

    @N("java/lang/Float", "toString", "(F)Ljava/lang/String;")
    def _(vm, a):
        return java_float_to_string(a[0])

    @N("java/lang/Float", "toString", "()Ljava/lang/String;")
    def _(vm, a):
        return java_float_to_string(a[0].p)

    @N("java/lang/Float", "hashCode")
    def _(vm, a):
        return float_to_raw_int_bits(a[0].p)

    @N("java/lang/Double", "doubleToRawLongBits")
    def _(vm, a):
        return struct.unpack("<q", struct.pack("<d", a[0]))[0]

    @N("java/lang/Double", "doubleToLongBits")
    def _(vm, a):
        return 0x7FF8000000000000 if a[0] != a[0] else struct.unpack("<q", struct.pack("<d", a[0]))[0]

    @N("java/lang/Double", "longBitsToDouble")
    def _(vm, a):
        return struct.unpack("<d", struct.pack("<q", a[0]))[0]

    @N("java/lang/Double", "isNaN")
    def _(vm, a):
        v = a[-1] if not isinstance(a[-1], JObject) else a[-1].p
        return 1 if v != v else 0

    @N("java/lang/Double", "isInfinite")
    def _(vm, a):
        v = a[-1] if not isinstance(a[-1], JObject) else a[-1].p
        return 1 if math.isinf(v) else 0

    @N("java/lang/Double", "compare")
    def _(vm, a):
        return _fcompare(a[0], a[1])

    @N("java/lang/Double", "valueOf", "(D)Ljava/lang/Double;")
    def _(vm, a):
        return box(vm, "java/lang/Double", a[0])

    @N("java/lang/Double", "toString", "(D)Ljava/lang/String;")
    def _(vm, a):
        return java_double_to_string(a[0])

    for cn, lo in (("java/lang/Integer", 32), ("java/lang/Long", 64), ("java/lang/Short", 16), ("java/lang/Byte", 8)):
        @N(cn, "valueOf", "(%s)L%s;" % ({32: "I", 64: "J", 16: "S", 8: "B"}[lo], cn))
        def _(vm, a, cn=cn):
            return box(vm, cn, a[0])

        @N(cn, "<init>")
        def _(vm, a):
            a[0].p = a[1] if not isinstance(a[1], str) else int(a[1])

        @N(cn, "intValue")
        def _(vm, a):
            return i32(a[0].p)

        @N(cn, "longValue")
        def _(vm, a):
            return a[0].p

        @N(cn, "floatValue")
        def _(vm, a):
            return f32(float(a[0].p))

        @N(cn, "doubleValue")
        def _(vm, a):
            return float(a[0].p)

        @N(cn, "hashCode")
        def _(vm, a, lo=lo):
            v = a[0].p
            return i32(v ^ ((v & 0xFFFFFFFFFFFFFFFF) >> 32)) if lo == 64 else v

        @N(cn, "equals")
        def _(vm, a):
            o = a[1]
            return 1 if isinstance(o, JObject) and o.cls is a[0].cls and o.p == a[0].p else 0

        @N(cn, "compareTo")
        def _(vm, a):
            return (a[0].p > a[1].p) - (a[0].p < a[1].p)

        @N(cn, "toString", "()Ljava/lang/String;")
        def _(vm, a):
            return str(a[0].p)

        @N(cn, "compare")
        def _(vm, a):
            return (a[0] > a[1]) - (a[0] < a[1])

    for cn in ("java/lang/Float", "java/lang/Double"):
        @N(cn, "floatValue")
        def _(vm, a):
            return f32(a[0].p)

        @N(cn, "doubleValue")
        def _(vm, a):
            return a[0].p

        @N(cn, "intValue")
        def _(vm, a):
            from minijvm import _f2i
            return _f2i(a[0].p, 31)

        @N(cn, "equals")
        def _(vm, a):
            o = a[1]
            return 1 if isinstance(o, JObject) and o.cls is a[0].cls and _fcompare(o.p, a[0].p) == 0 else 0

    @N("java/lang/Integer", "toString", "(I)Ljava/lang/String;")
    def _(vm, a):
        return str(a[0])

    @N("java/lang/Integer", "toHexString")
    def _(vm, a):
        return "%x" % (a[0] & 0xFFFFFFFF)

    @N("java/lang/Integer", "parseInt")
    def _(vm, a):
        try:
            return i32(int(a[0], a[1] if len(a) > 1 else 10))
        except ValueError:
            vm.throw("java/lang/NumberFormatException", a[0])

    @N("java/lang/Integer", "bitCount")
    def _(vm, a):
        return bin(a[0] & 0xFFFFFFFF).count("1")

    @N("java/lang/Integer", "numberOfLeadingZeros")
    def _(vm, a):
        return 32 - (a[0] & 0xFFFFFFFF).bit_length()

    @N("java/lang/Integer", "numberOfTrailingZeros")
    def _(vm, a):
        v = a[0] & 0xFFFFFFFF
        return 32 if v == 0 else (v & -v).bit_length() - 1

    @N("java/lang/Integer", "highestOneBit")
    def _(vm, a):
        v = a[0] & 0xFFFFFFFF
        return 0 if v == 0 else i32(1 << (v.bit_length() - 1))

    @N("java/lang/Integer", "rotateLeft")
    def _(vm, a):
        v, s = a[0] & 0xFFFFFFFF, a[1] & 31
        return i32((v << s) | (v >> (32 - s))) if s else i32(v)

    @N("java/lang/Integer", "reverseBytes")
    def _(vm, a):
        return struct.unpack("<i", struct.pack(">i", a[0]))[0]

    @N("java/lang/Long", "toString", "(J)Ljava/lang/String;")
    def _(vm, a):
        return str(a[0])

    @N("java/lang/Long", "bitCount")
    def _(vm, a):
        return bin(a[0] & 0xFFFFFFFFFFFFFFFF).count("1")

    @N("java/lang/Long", "numberOfLeadingZeros")
    def _(vm, a):
        return 64 - (a[0] & 0xFFFFFFFFFFFFFFFF).bit_length()

    @N("java/lang/Long", "numberOfTrailingZeros")
    def _(vm, a):
        v = a[0] & 0xFFFFFFFFFFFFFFFF
        return 64 if v == 0 else (v & -v).bit_length() - 1

    @N("java/lang/Long", "parseLong")
    def _(vm, a):
        try:
            return i64(int(a[0]))
        except ValueError:
            vm.throw("java/lang/NumberFormatException", a[0])

    @N("java/lang/Long", "rotateLeft")
    def _(vm, a):
        v, s = a[0] & 0xFFFFFFFFFFFFFFFF, a[1] & 63
        return i64((v << s) | (v >> (64 - s))) if s else i64(v)

    @N("java/lang/Boolean", "valueOf", "(Z)Ljava/lang/Boolean;")
    def _(vm, a):
        return box(vm, "java/lang/Boolean", 1 if a[0] else 0)

    @N("java/lang/Boolean", "booleanValue")
    def _(vm, a):
        return a[0].p

    @N("java/lang/Boolean", "parseBoolean")
    def _(vm, a):
        return 1 if a[0] is not None and a[0].lower() == "true" else 0

    @N("java/lang/Boolean", "getBoolean")
    def _(vm, a):
        return 0

    @N("java/lang/Boolean", "<clinit>")
    def _(vm, a):
        c = vm.load("java/lang/Boolean")
        for n, v in (("TRUE", 1), ("FALSE", 0)):
            c.fields[n] = (0x0019, "Ljava/lang/Boolean;")
            c.statics[n] = box(vm, "java/lang/Boolean", v)
    vm.load("java/lang/Boolean").init_state = 0

    @N("java/lang/Boolean", "equals")
    def _(vm, a):
        return 1 if isinstance(a[1], JObject) and a[1].cls is a[0].cls and a[1].p == a[0].p else 0

    @N("java/lang/Boolean", "hashCode")
    def _(vm, a):
        return 1231 if a[0].p else 1237

    @N("java/lang/Character", "toUpperCase")
    def _(vm, a):
        return ord(chr(a[0]).upper()) if len(chr(a[0]).upper()) == 1 else a[0]

    @N("java/lang/Character", "toLowerCase")
    def _(vm, a):
        return ord(chr(a[0]).lower())

    @N("java/lang/Character", "isWhitespace")
    def _(vm, a):
        return 1 if a[0] in (9, 10, 11, 12, 13, 28, 29, 30, 31, 32) else 0

    @N("java/lang/Character", "isDigit")
    def _(vm, a):
        return 1 if chr(a[0]).isdigit() else 0

    @N("java/lang/Character", "isLetter")
    def _(vm, a):
        return 1 if chr(a[0]).isalpha() else 0

    # ------------------------------------------------------------------ String / StringBuilder
    S = "java/lang/String"

    @N(S, "length")
    def _(vm, a):
        return len(a[0])

    @N(S, "charAt")
    def _(vm, a):
        if a[1] < 0 or a[1] >= len(a[0]):
            vm.throw("java/lang/StringIndexOutOfBoundsException", str(a[1]))
        return ord(a[0][a[1]])

    @N(S, "equals")
    def _(vm, a):
        return 1 if isinstance(a[1], str) and a[0] == a[1] else 0

    @N(S, "equalsIgnoreCase")
    def _(vm, a):
        return 1 if isinstance(a[1], str) and a[0].lower() == a[1].lower() else 0

    @N(S, "hashCode")
    def _(vm, a):
        return string_hash(a[0])

    @N(S, "compareTo")
    def _(vm, a):
        x, y = a[0], a[1]
        for c1, c2 in zip(x, y):
            if c1 != c2:
                return ord(c1) - ord(c2)
        return len(x) - len(y)

    @N(S, "toString")
    def _(vm, a):
        return a[0]

    @N(S, "intern")
    def _(vm, a):
        return a[0]

    @N(S, "isEmpty")
    def _(vm, a):
        return 1 if not a[0] else 0

    @N(S, "toUpperCase")
    def _(vm, a):
        return a[0].upper()

    @N(S, "toLowerCase")
    def _(vm, a):
        return a[0].lower()

    @N(S, "trim")
    def _(vm, a):
        s = a[0]
        i, j = 0, len(s)
        while i < j and s[i] <= " ":
            i += 1
        while j > i and s[j - 1] <= " ":
            j -= 1
        return s[i:j]

    @N(S, "startsWith")
    def _(vm, a):
        return 1 if a[0].startswith(a[1], a[2] if len(a) > 2 else 0) else 0

    @N(S, "endsWith")
    def _(vm, a):
        return 1 if a[0].endswith(a[1]) else 0

    @N(S, "indexOf")
    def _(vm, a):
        n = a[1] if isinstance(a[1], str) else chr(a[1])
        return a[0].find(n, a[2] if len(a) > 2 else 0)

    @N(S, "lastIndexOf")
    def _(vm, a):
        n = a[1] if isinstance(a[1], str) else chr(a[1])
        return a[0].rfind(n)

    @N(S, "contains")
    def _(vm, a):
        return 1 if jstr(vm, a[1]) in a[0] else 0

    @N(S, "substring")
    def _(vm, a):
        e = a[2] if len(a) > 2 else len(a[0])
        if a[1] < 0 or e > len(a[0]) or a[1] > e:
            vm.throw("java/lang/StringIndexOutOfBoundsException", "%d..%d" % (a[1], e))
        return a[0][a[1]:e]

    @N(S, "concat")
    def _(vm, a):
        return a[0] + a[1]

    @N(S, "replace")
    def _(vm, a):
        x = a[1] if isinstance(a[1], str) else chr(a[1])
        y = a[2] if isinstance(a[2], str) else chr(a[2])
        return a[0].replace(x, y)

    @N(S, "toCharArray")
    def _(vm, a):
        return JArray("C", [ord(c) for c in a[0]])

    @N(S, "getBytes")
    def _(vm, a):
        return JArray("B", [b - 256 if b > 127 else b for b in a[0].encode("utf-8")])

    @N(S, "valueOf")
    def _(vm, a):
        x = a[0]
        if isinstance(x, JArray):
            return "".join(chr(c) for c in x.a)
        if isinstance(x, float):
            return java_double_to_string(x)
        if isinstance(x, int):
            return str(x)
        return jstr(vm, x)

    @N(S, "valueOf", "(C)Ljava/lang/String;")
    def _(vm, a):
        return chr(a[0])

    @N(S, "valueOf", "(F)Ljava/lang/String;")
    def _(vm, a):
        return java_float_to_string(a[0])

    @N(S, "valueOf", "(Z)Ljava/lang/String;")
    def _(vm, a):
        return "true" if a[0] else "false"

    @N(S, "format")
    def _(vm, a):
        return a[-2] if isinstance(a[-2], str) else "<format>"

    # `new String(...)`: Strings are immutable Python strs, so `new` + `<init>` cannot fill the object in place;
    # the object created by `new` becomes a mutable holder whose payload is the str (handled by unwrap below).
    for sb in ("java/lang/StringBuilder", "java/lang/StringBuffer"):
        @N(sb, "<init>")
        def _(vm, a):
            a[0].p = [a[1]] if len(a) > 1 and isinstance(a[1], str) else []

        @N(sb, "append")
        def _(vm, a):
            x = a[1]
            if isinstance(x, JArray):
                if len(a) > 2:
                    x = "".join(chr(c) for c in x.a[a[2]:a[2] + a[3]])
                else:
                    x = "".join(chr(c) for c in x.a)
            elif isinstance(x, float):
                x = java_double_to_string(x)
            elif isinstance(x, int):
                x = str(x)
            else:
                x = jstr(vm, x)
            a[0].p.append(x)
            return a[0]

        @N(sb, "append", "(C)L%s;" % sb)
        def _(vm, a):
            a[0].p.append(chr(a[1]))
            return a[0]

        @N(sb, "append", "(F)L%s;" % sb)
        def _(vm, a):
            a[0].p.append(java_float_to_string(a[1]))
            return a[0]

        @N(sb, "append", "(Z)L%s;" % sb)
        def _(vm, a):
            a[0].p.append("true" if a[1] else "false")
            return a[0]

        @N(sb, "toString")
        def _(vm, a):
            s = "".join(a[0].p)
            a[0].p = [s]
            return s

        @N(sb, "length")
        def _(vm, a):
            return len("".join(a[0].p))

        @N(sb, "setLength")
        def _(vm, a):
            a[0].p = ["".join(a[0].p)[:a[1]]]

        @N(sb, "charAt")
        def _(vm, a):
            return ord("".join(a[0].p)[a[1]])

    # ------------------------------------------------------------------ Arrays / Collections / Objects
    A = "java/util/Arrays"

    @N(A, "fill")
    def _(vm, a):
        if len(a) == 2:
            arr, v = a
            arr.a[:] = [v] * len(arr.a)
        else:
            arr, lo, hi, v = a
            arr.a[lo:hi] = [v] * (hi - lo)

    @N(A, "copyOf")
    def _(vm, a):
        arr, n = a[0], a[1]
        z = None if arr.etype[0] in "L[" else (0.0 if arr.etype in "FD" else 0)
        return JArray(arr.etype, list(arr.a[:n]) + [z] * max(0, n - len(arr.a)))

    @N(A, "copyOfRange")
    def _(vm, a):
        arr, lo, hi = a[0], a[1], a[2]
        z = None if arr.etype[0] in "L[" else (0.0 if arr.etype in "FD" else 0)
        out = list(arr.a[lo:hi])
        return JArray(arr.etype, out + [z] * (hi - lo - len(out)))

    @N(A, "asList")
    def _(vm, a):
        return mk(vm, "java/util/Arrays$ArrayList", list(a[0].a))

    @N(A, "equals")
    def _(vm, a):
        x, y = a
        if x is y:
            return 1
        if x is None or y is None or len(x.a) != len(y.a):
            return 0
        if x.etype[0] in "L[":
            return 1 if all(jequals(vm, p, q) for p, q in zip(x.a, y.a)) else 0
        return 1 if x.a == y.a else 0

    @N(A, "hashCode")
    def _(vm, a):
        x = a[0]
        if x is None:
            return 0
        h = 1
        for e in x.a:
            if x.etype[0] in "L[":
                eh = jhash(vm, e)
            elif x.etype == "J":
                eh = i32(e ^ ((e & 0xFFFFFFFFFFFFFFFF) >> 32))
            elif x.etype == "F":
                eh = float_to_raw_int_bits(e)
            elif x.etype == "Z":
                eh = 1231 if e else 1237
            else:
                eh = e
            h = (31 * h + eh) & 0xFFFFFFFF
        return i32(h)

    @N(A, "toString")
    def _(vm, a):
        if a[0] is None:
            return "null"
        return "[" + ", ".join(jstr(vm, x) if not isinstance(x, (int, float)) else str(x) for x in a[0].a) + "]"

    def _sort_key(vm, cmp):
        import functools
        if cmp is None:
            return functools.cmp_to_key(lambda p, q: vm.call_virtual(p, "compareTo", "(Ljava/lang/Object;)I", q))
        return functools.cmp_to_key(lambda p, q: vm.call_virtual(cmp, "compare", "(Ljava/lang/Object;Ljava/lang/Object;)I", p, q))

    @N(A, "sort")
    def _(vm, a):
        arr = a[0]
        if arr.etype[0] not in "L[":
            if len(a) == 3:
                arr.a[a[1]:a[2]] = sorted(arr.a[a[1]:a[2]])
            else:
                arr.a.sort()
            return
        if len(a) >= 3 and isinstance(a[1], int):
            cmp = a[3] if len(a) > 3 else None
            arr.a[a[1]:a[2]] = sorted(arr.a[a[1]:a[2]], key=_sort_key(vm, cmp))
        else:
            cmp = a[1] if len(a) > 1 else None
            arr.a.sort(key=_sort_key(vm, cmp))  # Python's sort is stable like Java's object sort

    C = "java/util/Collections"

    @N(C, "unmodifiableList")
    @N(C, "unmodifiableCollection")
    @N(C, "unmodifiableSet")
    @N(C, "unmodifiableMap")
    @N(C, "synchronizedSet")
    @N(C, "synchronizedMap")
    @N(C, "synchronizedList")
    def _(vm, a):
        return a[0]

    @N(C, "emptyList")
    def _(vm, a):
        return mk(vm, "java/util/ArrayList", [])

    @N(C, "emptySet")
    def _(vm, a):
        return mk(vm, "java/util/HashSet", [])

    @N(C, "emptyMap")
    def _(vm, a):
        return mk(vm, "java/util/HashMap", [])

    @N(C, "emptyIterator")
    def _(vm, a):
        return make_iter(vm, [])

    @N(C, "singletonList")
    @N(C, "singleton")
    def _(vm, a):
        return mk(vm, "java/util/ArrayList", [a[0]])

    @N(C, "singletonMap")
    def _(vm, a):
        return mk(vm, "java/util/HashMap", [[a[0], a[1]]])

    @N(C, "newSetFromMap")
    def _(vm, a):
        return mk(vm, "java/util/HashSet", [])

    @N(C, "sort")
    def _(vm, a):
        a[0].p.sort(key=_sort_key(vm, a[1] if len(a) > 1 else None))

    @N(C, "addAll")
    def _(vm, a):
        for x in a[1].a:
            vm.call_virtual(a[0], "add", "(Ljava/lang/Object;)Z", x)
        return 1

    @N(C, "reverse")
    def _(vm, a):
        a[0].p.reverse()

    @N(C, "swap")
    def _(vm, a):
        l = a[0].p
        l[a[1]], l[a[2]] = l[a[2]], l[a[1]]

    @N("java/util/Objects", "equals")
    def _(vm, a):
        return 1 if jequals(vm, a[0], a[1]) else 0

    @N("java/util/Objects", "hashCode")
    def _(vm, a):
        return jhash(vm, a[0])

    @N("java/util/Objects", "requireNonNull")
    def _(vm, a):
        if a[0] is None:
            vm.throw("java/lang/NullPointerException", a[1] if len(a) > 1 and isinstance(a[1], str) else None)
        return a[0]

    @N("java/util/Objects", "hash")
    def _(vm, a):
        h = 1
        for e in a[0].a:
            h = (31 * h + jhash(vm, e)) & 0xFFFFFFFF
        return i32(h)

    @N("java/util/Objects", "toString")
    def _(vm, a):
        return jstr(vm, a[0])

    # ------------------------------------------------------------------ lists
    for L in ("java/util/ArrayList", "java/util/AbstractList", "java/util/AbstractCollection"):
        @N(L, "<init>")
        def _(vm, a):
            if len(a) > 1 and isinstance(a[1], JObject):
                a[0].p = list(_iterate(vm, a[1]))
            else:
                a[0].p = []

    L = "java/util/ArrayList"

    @N(L, "add")
    def _(vm, a):
        if len(a) == 3:
            a[0].p.insert(a[1], a[2])
            return None
        a[0].p.append(a[1])
        return 1

    @N(L, "addAll")
    def _(vm, a):
        xs = list(_iterate(vm, a[-1]))
        if len(a) == 3:
            a[0].p[a[1]:a[1]] = xs
        else:
            a[0].p.extend(xs)
        return 1 if xs else 0

    @N(L, "get")
    def _(vm, a):
        if a[1] < 0 or a[1] >= len(a[0].p):
            vm.throw("java/lang/IndexOutOfBoundsException", "Index: %d, Size: %d" % (a[1], len(a[0].p)))
        return a[0].p[a[1]]

    @N(L, "set")
    def _(vm, a):
        old = a[0].p[a[1]]
        a[0].p[a[1]] = a[2]
        return old

    @N(L, "size")
    def _(vm, a):
        return len(a[0].p)

    @N(L, "isEmpty")
    def _(vm, a):
        return 0 if a[0].p else 1

    @N(L, "clear")
    def _(vm, a):
        del a[0].p[:]

    @N(L, "remove", "(I)Ljava/lang/Object;")
    def _(vm, a):
        return a[0].p.pop(a[1])

    @N(L, "remove", "(Ljava/lang/Object;)Z")
    def _(vm, a):
        for i, x in enumerate(a[0].p):
            if jequals(vm, a[1], x):
                del a[0].p[i]
                return 1
        return 0

    @N(L, "contains")
    def _(vm, a):
        return 1 if any(jequals(vm, a[1], x) for x in a[0].p) else 0

    @N(L, "indexOf")
    def _(vm, a):
        for i, x in enumerate(a[0].p):
            if jequals(vm, a[1], x):
                return i
        return -1

    @N(L, "iterator")
    @N(L, "listIterator")
    def _(vm, a):
        it = make_iter(vm, a[0].p)
        it.p[2] = a[0]
        return it

    @N(L, "toArray")
    def _(vm, a):
        if len(a) > 1:
            return JArray(a[1].etype, list(a[0].p)) if len(a[1].a) < len(a[0].p) else _fill(a[1], a[0].p)
        return JArray("Ljava/lang/Object;", list(a[0].p))

    @N(L, "subList")
    def _(vm, a):
        return mk(vm, "java/util/ArrayList", a[0].p[a[1]:a[2]])

    @N(L, "equals")
    def _(vm, a):
        o = a[1]
        if not isinstance(o, JObject) or not isinstance(o.p, list) or "java/util/List" not in o.cls.supers():
            return 0
        return 1 if len(o.p) == len(a[0].p) and all(jequals(vm, p, q) for p, q in zip(a[0].p, o.p)) else 0

    @N(L, "hashCode")
    def _(vm, a):
        h = 1
        for e in a[0].p:
            h = (31 * h + jhash(vm, e)) & 0xFFFFFFFF
        return i32(h)

    @N(L, "toString")
    def _(vm, a):
        return "[" + ", ".join(jstr(vm, x) for x in a[0].p) + "]"

    @N(L, "ensureCapacity")
    @N(L, "trimToSize")
    def _(vm, a):
        return None

    # LinkedList / Queue extras
    @N("java/util/LinkedList", "poll")
    @N("java/util/LinkedList", "pollFirst")
    def _(vm, a):
        return a[0].p.pop(0) if a[0].p else None

    @N("java/util/LinkedList", "peek")
    def _(vm, a):
        return a[0].p[0] if a[0].p else None

    @N("java/util/LinkedList", "offer")
    @N("java/util/LinkedList", "addLast")
    def _(vm, a):
        a[0].p.append(a[1])
        return 1

    @N("java/util/LinkedList", "removeFirst")
    def _(vm, a):
        return a[0].p.pop(0)

    @N("java/util/LinkedList", "getLast")
    def _(vm, a):
        return a[0].p[-1]

    # ------------------------------------------------------------------ iterator / entries
    I = "java/util/PyIterator"

    @N(I, "hasNext")
    def _(vm, a):
        return 1 if a[0].p[1] < len(a[0].p[0]) else 0

    @N(I, "next")
    def _(vm, a):
        p = a[0].p
        if p[1] >= len(p[0]):
            vm.throw("java/util/NoSuchElementException")
        p[1] += 1
        return p[0][p[1] - 1]

    @N(I, "remove")
    def _(vm, a):
        p = a[0].p
        owner = p[2]
        if owner is None:
            vm.throw("java/lang/UnsupportedOperationException", "remove")
        x = p[0][p[1] - 1]
        if isinstance(owner.p, list):
            for i, y in enumerate(owner.p):
                if y is x:
                    del owner.p[i]
                    break

    E = "java/util/PyEntry"

    @N(E, "getKey")
    def _(vm, a):
        return a[0].p[0]

    @N(E, "getValue")
    def _(vm, a):
        return a[0].p[1]

    @N(E, "setValue")
    def _(vm, a):
        old = a[0].p[1]
        a[0].p[1] = a[1]
        return old

    # ------------------------------------------------------------------ maps (list of [k, v]; identity variant)
    M = "java/util/HashMap"

    def _find(vm, mp, k):
        ident = mp.cls.name == "java/util/IdentityHashMap"
        for e in mp.p:
            if (e[0] is k) if ident else jequals(vm, k, e[0]):
                return e
        return None

    for MM in ("java/util/HashMap", "java/util/AbstractMap"):
        @N(MM, "<init>")
        def _(vm, a):
            a[0].p = []
            if len(a) > 1 and isinstance(a[1], JObject) and isinstance(a[1].p, list) and "java/util/Map" in a[1].cls.supers():
                a[0].p = [list(e) for e in a[1].p]

    @N(M, "put")
    def _(vm, a):
        e = _find(vm, a[0], a[1])
        if e is not None:
            old = e[1]
            e[1] = a[2]
            return old
        a[0].p.append([a[1], a[2]])
        return None

    @N(M, "putAll")
    def _(vm, a):
        for k, v in a[1].p:
            e = _find(vm, a[0], k)
            if e is not None:
                e[1] = v
            else:
                a[0].p.append([k, v])

    @N(M, "get")
    def _(vm, a):
        e = _find(vm, a[0], a[1])
        return None if e is None else e[1]

    @N(M, "containsKey")
    def _(vm, a):
        return 0 if _find(vm, a[0], a[1]) is None else 1

    @N(M, "remove")
    def _(vm, a):
        e = _find(vm, a[0], a[1])
        if e is None:
            return None
        a[0].p.remove(e)
        return e[1]

    @N(M, "size")
    def _(vm, a):
        return len(a[0].p)

    @N(M, "isEmpty")
    def _(vm, a):
        return 0 if a[0].p else 1

    @N(M, "clear")
    def _(vm, a):
        del a[0].p[:]

    @N(M, "keySet")
    def _(vm, a):
        return mk(vm, "java/util/LinkedHashSet", [e[0] for e in a[0].p])

    @N(M, "values")
    def _(vm, a):
        return mk(vm, "java/util/ArrayList", [e[1] for e in a[0].p])

    @N(M, "entrySet")
    def _(vm, a):
        return mk(vm, "java/util/LinkedHashSet", [mk(vm, "java/util/PyEntry", e) for e in a[0].p])

    # ------------------------------------------------------------------ sets (list)
    H = "java/util/HashSet"

    @N(H, "<init>")
    @N("java/util/AbstractSet", "<init>")
    def _(vm, a):
        a[0].p = []
        if len(a) > 1 and isinstance(a[1], JObject):
            for x in _iterate(vm, a[1]):
                if not any(jequals(vm, x, y) for y in a[0].p):
                    a[0].p.append(x)

    @N(H, "add")
    def _(vm, a):
        if any(jequals(vm, a[1], y) for y in a[0].p):
            return 0
        a[0].p.append(a[1])
        return 1

    @N(H, "addAll")
    def _(vm, a):
        ch = 0
        for x in _iterate(vm, a[1]):
            if not any(jequals(vm, x, y) for y in a[0].p):
                a[0].p.append(x)
                ch = 1
        return ch

    @N(H, "contains")
    def _(vm, a):
        return 1 if any(jequals(vm, a[1], y) for y in a[0].p) else 0

    @N(H, "remove")
    def _(vm, a):
        for i, y in enumerate(a[0].p):
            if jequals(vm, a[1], y):
                del a[0].p[i]
                return 1
        return 0

    for nm in ("size", "isEmpty", "clear", "iterator", "toArray", "toString"):
        vm.natives[(H, nm, None)] = vm.natives[(L, nm, None)]

    # ------------------------------------------------------------------ atomics / misc used by reader plumbing
    for cn in ("java/util/concurrent/atomic/AtomicInteger", "java/util/concurrent/atomic/AtomicLong"):
        @N(cn, "<init>")
        def _(vm, a):
            a[0].p = a[1] if len(a) > 1 else 0

        @N(cn, "get")
        @N(cn, "intValue")
        @N(cn, "longValue")
        def _(vm, a):
            return a[0].p

        @N(cn, "set")
        def _(vm, a):
            a[0].p = a[1]

        @N(cn, "incrementAndGet")
        def _(vm, a):
            a[0].p += 1
            return a[0].p

        @N(cn, "decrementAndGet")
        def _(vm, a):
            a[0].p -= 1
            return a[0].p

        @N(cn, "getAndIncrement")
        def _(vm, a):
            a[0].p += 1
            return a[0].p - 1

        @N(cn, "addAndGet")
        def _(vm, a):
            a[0].p += a[1]
            return a[0].p

        @N(cn, "compareAndSet")
        def _(vm, a):
            if a[0].p == a[1]:
                a[0].p = a[2]
                return 1
            return 0

    @N("java/util/concurrent/atomic/AtomicBoolean", "<init>")
    def _(vm, a):
        a[0].p = a[1] if len(a) > 1 else 0

    @N("java/util/concurrent/atomic/AtomicBoolean", "get")
    def _(vm, a):
        return a[0].p

    @N("java/util/concurrent/atomic/AtomicBoolean", "set")
    def _(vm, a):
        a[0].p = a[1]

    @N("java/util/concurrent/atomic/AtomicBoolean", "compareAndSet")
    def _(vm, a):
        if a[0].p == a[1]:
            a[0].p = a[2]
            return 1
        return 0

    @N("java/lang/ThreadLocal", "<init>")
    def _(vm, a):
        a[0].p = [False, None]

    @N("java/lang/ThreadLocal", "get")
    def _(vm, a):
        if not a[0].p[0]:
            a[0].p = [True, vm.call_virtual(a[0], "initialValue", "()Ljava/lang/Object;")]
        return a[0].p[1]

    @N("java/lang/ThreadLocal", "initialValue")
    def _(vm, a):
        return None

    @N("java/lang/ThreadLocal", "set")
    def _(vm, a):
        a[0].p = [True, a[1]]

    # ------------------------------------------------------------------ java.io readers (fasta-utils)
    @N("java/io/FileInputStream", "<init>")
    def _(vm, a):
        x = a[1]
        path = x if isinstance(x, str) else x.p
        try:
            a[0].p = open(path, "rb")
        except OSError as e:
            vm.throw("java/io/FileNotFoundException", str(e))

    @N("java/io/File", "<init>")
    def _(vm, a):
        a[0].p = a[1] if len(a) == 2 else (a[1] if isinstance(a[1], str) else a[1].p) + "/" + a[2]

    @N("java/io/File", "getAbsolutePath")
    @N("java/io/File", "getPath")
    @N("java/io/File", "toString")
    def _(vm, a):
        return a[0].p

    @N("java/io/File", "exists")
    def _(vm, a):
        import os
        return 1 if os.path.exists(a[0].p) else 0

    @N("java/io/InputStreamReader", "<init>")
    def _(vm, a):
        src = a[1].p
        cs = a[2] if len(a) > 2 else "US-ASCII"
        if isinstance(cs, JObject):
            cs = cs.p
        data = src.read()
        # US-ASCII decoding in Java maps bytes ≥ 0x80 to U+FFFD
        if cs.upper() in ("US-ASCII", "ASCII"):
            text = "".join(chr(b) if b < 128 else "�" for b in data)
        else:
            text = data.decode(cs, "replace")
        a[0].p = [text, 0]

    @N("java/io/StringReader", "<init>")
    def _(vm, a):
        a[0].p = [a[1], 0]

    @N("java/io/BufferedReader", "<init>")
    def _(vm, a):
        a[0].p = a[1].p

    @N("java/io/BufferedReader", "readLine")
    def _(vm, a):
        p = a[0].p
        text, pos = p
        n = len(text)
        if pos >= n:
            return None
        i = pos
        while i < n and text[i] != "\n" and text[i] != "\r":
            i += 1
        line = text[pos:i]
        if i < n:
            if text[i] == "\r" and i + 1 < n and text[i + 1] == "\n":
                i += 2
            else:
                i += 1
        p[1] = i
        return line

    for cn in ("java/io/Reader", "java/io/BufferedReader", "java/io/InputStreamReader", "java/io/StringReader"):
        @N(cn, "read", "([CII)I")
        def _(vm, a):
            p = a[0].p
            text, pos = p
            if pos >= len(text):
                return -1
            n = min(a[3], len(text) - pos)
            a[1].a[a[2]:a[2] + n] = [ord(c) for c in text[pos:pos + n]]
            p[1] = pos + n
            return n

        @N(cn, "read", "()I")
        def _(vm, a):
            p = a[0].p
            if p[1] >= len(p[0]):
                return -1
            p[1] += 1
            return ord(p[0][p[1] - 1])

        @N(cn, "close")
        def _(vm, a):
            return None

    @N("java/io/FileInputStream", "close")
    @N("java/io/InputStream", "close")
    def _(vm, a):
        if a[0].p is not None and hasattr(a[0].p, "close"):
            a[0].p.close()

    @N("java/nio/charset/Charset", "forName")
    def _(vm, a):
        return mk(vm, "java/nio/charset/Charset", a[0])

    @N("java/nio/charset/StandardCharsets", "<clinit>")
    def _(vm, a):
        c = vm.load("java/nio/charset/StandardCharsets")
        for n, v in (("US_ASCII", "US-ASCII"), ("UTF_8", "UTF-8"), ("ISO_8859_1", "ISO-8859-1")):
            c.fields[n] = (0x0019, "Ljava/nio/charset/Charset;")
            c.statics[n] = mk(vm, "java/nio/charset/Charset", v)
    vm.load("java/nio/charset/StandardCharsets").init_state = 0

    # java.util.regex for the two header patterns of fasta-utils (`^(\S+)\s*$`, `^(\S+)\s+(\S+.*)$`): Python's `re`
    # with re.ASCII has the same meaning for \S, \s ([ \t\n\x0B\f\r]), ., ^, $, +, * on these inputs (no newlines)
    @N("java/util/regex/Pattern", "compile")
    def _(vm, a):
        import re
        return mk(vm, "java/util/regex/Pattern", re.compile(a[0], re.ASCII))

    @N("java/util/regex/Pattern", "matcher")
    def _(vm, a):
        return mk(vm, "java/util/regex/Matcher", [a[0].p, jstr(vm, a[1]), None])

    @N("java/util/regex/Matcher", "matches")
    def _(vm, a):
        p = a[0].p
        p[2] = p[0].fullmatch(p[1])
        return 1 if p[2] is not None else 0

    @N("java/util/regex/Matcher", "find")
    def _(vm, a):
        p = a[0].p
        p[2] = p[0].search(p[1], p[2].end() if p[2] is not None else 0)
        return 1 if p[2] is not None else 0

    @N("java/util/regex/Matcher", "group")
    def _(vm, a):
        return a[0].p[2].group(a[1] if len(a) > 1 else 0)

    @N(S, "split")
    def _(vm, a):
        # String#split(regex) for literal (metacharacter-free) patterns; trailing empty strings are removed
        import re
        pat = re.sub(r"\\c([A-Z])", lambda m: "\\x%02x" % (ord(m.group(1)) - 64), a[1])  # \cA = Ctrl-A
        if any(ch in pat.replace("\\x", "") for ch in "\\[](){}.*+?^$|"):
            raise RuntimeError("String.split with a regex pattern %r is not supported" % a[1])
        parts = re.split(pat, a[0]) if pat else list(a[0])
        while len(parts) > 1 and parts[-1] == "":
            parts.pop()
        return JArray("Ljava/lang/String;", parts)


def _fcompare(x, y):
    if x < y:
        return -1
    if x > y:
        return 1
    if x == y and x != 0.0:
        return 0
    # zeros and NaNs by bit pattern order: -0.0 < 0.0 < NaN
    kx = 2 if x != x else (0 if math.copysign(1, x) < 0 else 1)
    ky = 2 if y != y else (0 if math.copysign(1, y) < 0 else 1)
    return (kx > ky) - (kx < ky)


def _fill(arr, xs):
    arr.a[:len(xs)] = xs
    if len(arr.a) > len(xs):
        arr.a[len(xs)] = None
    return arr


def _iterate(vm, coll):
    """Iterate any Java Iterable (native list/set payloads directly, interpreted ones through iterator())."""
    if isinstance(coll.p, list) and coll.cls.native:
        if "java/util/Map" in coll.cls.supers():
            raise TypeError("iterate a map")
        return list(coll.p)
    it = vm.call_virtual(coll, "iterator", "()Ljava/util/Iterator;")
    out = []
    while vm.call_virtual(it, "hasNext", "()Z"):
        out.append(vm.call_virtual(it, "next", "()Ljava/lang/Object;"))
    return out
