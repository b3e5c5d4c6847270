#!/usr/bin/env python
"""Which kernels of two builds of libdce_b200.so differ in their SASS (instruction text, encodings ignored)?
Used to show that adding an option-gated template variant leaves the measured default kernels byte-identical.
    git stash; python -m deep_contact_estimator_b200.build --force; cp deep_contact_estimator_b200/libdce_b200.so /tmp/before.so
    git stash pop; python -m deep_contact_estimator_b200.build --force
    python tools/compare_sass.py /tmp/before.so deep_contact_estimator_b200/libdce_b200.so
Kernel names are compared after dropping defaulted trailing template arguments and dependent parameter types, so
`block1_kernel<false>` matches `block1_kernel<false, 0>`."""
import re
import subprocess
import sys


def kernels(path):
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, name, body = {}, None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name:
            t = re.sub(r"/\*[0-9a-fx ]+\*/", "", line).strip()
            if t:
                body.append(t)
    if name:
        out[name] = body
    return out


def demangle(names):
    res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    return dict(zip(names, res))


def canon(pretty):
    """`f<a, b, 0>(T)` -> ('f', (a, b, 0)) with trailing zeros / false dropped."""
    m = re.match(r"(?:void )?([\w:]+)<(.*?)>\(", pretty)
    if not m:
        return pretty.split("(")[0], ()
    args = [a.strip() for a in m.group(2).split(",")]
    args = [re.sub(r"^\((?:int|bool)\)", "", a) for a in args]
    args = ["0" if a == "false" else "1" if a == "true" else a for a in args]
    while args and args[-1] == "0":
        args.pop()
    return m.group(1), tuple(args)


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    da, db = demangle(list(a)), demangle(list(b))
    cb = {canon(db[k]): k for k in b}
    same = diff = missing = 0
    for k, body in a.items():
        kb = cb.get(canon(da[k]))
        if kb is None:
            missing += 1
            print("only in the first build:", da[k].split("(")[0])
        elif b[kb] == body:
            same += 1
        else:
            diff += 1
            print(f"DIFFERENT: {da[k].split('(')[0]}  ({len(body)} vs {len(b[kb])} instructions)")
    new = [db[k].split("(")[0] for k in b if canon(db[k]) not in {canon(da[x]) for x in a}]
    print(f"{same} kernels identical, {diff} different, {missing} missing; {len(new)} new in the second build")
    for n in new:
        print("  new:", n)
    return 1 if diff or missing else 0


if __name__ == "__main__":
    sys.exit(main())
