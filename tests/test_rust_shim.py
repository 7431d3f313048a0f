"""The Rust drop-in crates (rust/scz-sys, rust/dist-primitive-gpu, rust/hyperplonk-gpu) cannot be compiled in this image
(no cargo / rustc), so they are held to the C header mechanically:
  * rust/scz-sys/src/ffi.rs is regenerated from include/scz.h and compared (every prototype, nothing else);
  * the hand-written #[repr(C)] structs of scz-sys list the header's struct fields in the same order;
  * every function the shim crates call exists in the header, and every reference-named entry point of the hot path
     is present in the shim with the reference's argument names."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "rust")


def _read(*p):
    return open(os.path.join(*p)).read()


def test_ffi_block_is_generated_from_the_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_scz_sys.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    sys.path.insert(0, ROOT)
    from scz_b200 import binding
    ffi = set(re.findall(r"pub fn (scz_\w+)", _read(RUST, "scz-sys", "src", "ffi.rs")))
    assert ffi == set(binding.declared_symbols())


def _c_struct_fields(header, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        if "(*" in decl:                                   # function pointer member
            fields.append(re.search(r"\(\*\s*(\w+)\)", decl).group(1))
            continue
        names = decl.split(",")
        names[0] = names[0].split()[-1]
        fields += [n.strip().lstrip("*").strip() for n in names]
    return fields


def _rust_struct_fields(src, name):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, flags=re.S).group(1)
    return re.findall(r"pub (\w+)\s*:", body)


def test_repr_c_structs_match_the_header_field_by_field():
    header = _read(ROOT, "include", "scz.h")
    lib = _read(RUST, "scz-sys", "src", "lib.rs")
    for c_name, rs_name in (("scz_hp_pk", "SczHpPk"), ("scz_hp_item", "SczHpItem"), ("scz_local_pk", "SczLocalPk"),
                            ("scz_cperm_pk", "SczCpermPk"), ("scz_net_vtable", "SczNetVtable")):
        c = [f.lower() for f in _c_struct_fields(header, c_name)]
        r = _rust_struct_fields(lib, rs_name)
        assert c == r, (c_name, c, r)
    for const in re.findall(r"#define (SCZ_[A-Z0-9_]+) ", header):
        if const.startswith("SCZ_K_") or const == "SCZ_H":
            continue
        assert re.search(r"pub const %s\b" % const, lib), const


def test_shim_calls_only_declared_entry_points_and_keeps_the_reference_names():
    sys.path.insert(0, ROOT)
    from scz_b200 import binding
    declared = set(binding.declared_symbols())
    used = set()
    srcs = {}
    for crate in ("dist-primitive-gpu", "hyperplonk-gpu"):
        d = os.path.join(RUST, crate, "src")
        for f in os.listdir(d):
            srcs[f"{crate}/{f}"] = _read(d, f)
            used |= set(re.findall(r"\b(scz_[a-z0-9_]+)\s*\(", srcs[f"{crate}/{f}"]))
            used |= set(re.findall(r"\b(scz_[a-z0-9_]+_dev)\b", srcs[f"{crate}/{f}"]))
    assert used <= declared, sorted(used - declared)
    # reference name -> (file, argument names in the reference's order; the net handle may be an extra first argument)
    want = {
        "d_msm": ("dist-primitive-gpu/dmsm.rs", ["bases", "scalars", "pp", "net"]),                       # dmsm.rs:9-15
        "pss2ss": ("dist-primitive-gpu/unpack.rs", ["share", "pp", "net"]),                                # unpack.rs:72-77
        "degree_reduce": ("dist-primitive-gpu/degree_reduce.rs", ["shares", "pp", "net"]),                 # degree_reduce.rs:29-34
        "fix_variable": ("dist-primitive-gpu/mle.rs", ["evaluations", "points"]),                          # mle.rs:88-91
        "sumcheck": ("dist-primitive-gpu/dsumcheck.rs", ["evaluation", "challenge"]),                      # dsumcheck.rs:6
        "sumcheck_product": ("dist-primitive-gpu/dsumcheck.rs", ["evaluation_f", "evaluation_g", "challenge"]),
        "c_sumcheck": ("dist-primitive-gpu/dsumcheck.rs", ["shares", "challenge", "pp", "net"]),
        "c_sumcheck_product": ("dist-primitive-gpu/dsumcheck.rs", ["shares_f", "shares_g", "challenge", "pp", "net"]),
        "d_sumcheck": ("dist-primitive-gpu/dsumcheck.rs", ["partial_poly", "challenge", "net"]),
        "d_sumcheck_product": ("dist-primitive-gpu/dsumcheck.rs", ["partial_f", "partial_g", "challenge", "net"]),
        "acc_product": ("dist-primitive-gpu/dacc_product.rs", ["x"]),
        "d_acc_product": ("dist-primitive-gpu/dacc_product.rs", ["inputs", "net"]),
        "c_acc_product_and_share": ("dist-primitive-gpu/dacc_product.rs", ["shares", "masks", "unmask0", "unmask1", "unmask2", "pp", "net"]),
        "commit": ("dist-primitive-gpu/dpoly_comm.rs", ["peval"]),
        "c_commit": ("dist-primitive-gpu/dpoly_comm.rs", ["pevals", "pp", "net"]),
        "d_commit": ("dist-primitive-gpu/dpoly_comm.rs", ["peval", "net"]),
        "open": ("dist-primitive-gpu/dpoly_comm.rs", ["peval", "point"]),
        "c_open": ("dist-primitive-gpu/dpoly_comm.rs", ["peval", "point", "pp", "net"]),
        "d_open": ("dist-primitive-gpu/dpoly_comm.rs", ["peval", "point", "net"]),
        "dhyperplonk": ("hyperplonk-gpu/lib.rs", ["n", "pk", "pp", "net"]),                                # dhyperplonk.rs:159-165
        "dhyperplonk_data_parallel": ("hyperplonk-gpu/lib.rs", ["n", "pk", "pp", "net"]),
        "dpermcheck": ("hyperplonk-gpu/lib.rs", ["n", "pk", "pp", "net"]),
    }
    for name, (f, args) in want.items():
        m = re.search(r"pub (?:async )?fn %s\s*(?:<[^(]*>)?\s*\((.*?)\)\s*(?:->|\{)" % name, srcs[f], flags=re.S)
        assert m, (name, f)
        got = [a for a in re.findall(r"(?:^|,)\s*&?\s*(\w+)\s*:(?!:)", m.group(1)) if a not in ("self",)]
        got = [a.lstrip("_") for a in got if a.lstrip("_") != "sid"]
        if got and got[0] == "net" and (not args or args[0] != "net") and "net" not in args:
            got = got[1:]
        assert got == args, (name, got, args)
