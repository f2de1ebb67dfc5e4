#!/usr/bin/env python3
"""Extract the reference's known-answer tests into tests/golden/kats.json.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box).  The output is committed; tests read the JSON.

Every `#[test]` function of the corrector modules is literal data: byte
strings, a `Solid::new(k)`, `Tokenizer` loops that insert every k-mer of a
string, optional single `data.set(seq2bit(b"..."), true)` insertions, a
corrector constructor and `assert_eq!(expected, corrector.correct(input))`
lines.  This script interprets exactly those statement shapes and refuses
anything else, so a silent mis-parse cannot happen.

Sources (reference file:line of each test module):
  src/correct/mod.rs:166-181          helpers (alt_nucs KAT)
  src/correct/exist/one.rs:76-277     One
  src/correct/exist/two.rs:330-642    Two
  src/correct/graph.rs:88-318         Graph
  src/correct/greedy.rs:176-411       Greedy (3 are #[ignore])
  src/correct/gap_size.rs:111-258     GapSize
  src/set/pcon.rs:198-255             Pcon set
"""
import json
import re
import sys
from pathlib import Path

REF = Path("/root/reference/src")
FILES = [
    ("one", "correct/exist/one.rs"),
    ("two", "correct/exist/two.rs"),
    ("graph", "correct/graph.rs"),
    ("greedy", "correct/greedy.rs"),
    ("gap_size", "correct/gap_size.rs"),
]

BYTES = r'b"([A-Za-z\-]*)"'


def split_tests(src):
    """Yield (name, ignored, first_line, body) for each #[test] fn."""
    lines = src.split("\n")
    i = 0
    while i < len(lines):
        if lines[i].strip() == "#[test]":
            j = i + 1
            ignored = False
            if lines[j].strip() == "#[ignore]":
                ignored = True
                j += 1
            m = re.match(r"\s*fn (\w+)\(\)", lines[j])
            assert m, lines[j]
            name = m.group(1)
            depth = 0
            body = []
            k = j
            while True:
                depth += lines[k].count("{") - lines[k].count("}")
                body.append(lines[k])
                if depth == 0:
                    break
                k += 1
            yield name, ignored, j + 1, "\n".join(body[1:-1])
            i = k
        i += 1


def strip_comments(body):
    return "\n".join(re.sub(r"//.*$", "", l) for l in body.split("\n"))


def parse_test(module, name, body, statics):
    body = strip_comments(body)
    # join statements
    text = re.sub(r"\s+", " ", body)
    env = dict(statics)  # name -> bytes (str)
    k = None
    inserts_seq = []  # strings whose every k-mer is inserted
    inserts_kmer = []  # single k-mers inserted
    asserts = []
    corrector = None
    uses_get_solid = False

    stmts = [s.strip() for s in re.split(r";", text) if s.strip()]
    idx = 0
    while idx < len(stmts):
        s = stmts[idx]
        idx += 1
        m = re.fullmatch(r"let (\w+) = " + BYTES, s)
        if m:
            env[m.group(1)] = m.group(2)
            continue
        m = re.fullmatch(r"let (\w+) = filter\(" + BYTES + r"\)", s)
        if m:
            env[m.group(1)] = m.group(2).replace("-", "")
            continue
        m = re.fullmatch(
            r"let mut data(?:: pcon::solid::Solid)? = pcon::solid::Solid::new\((\d+)\)", s
        )
        if m:
            k = int(m.group(1))
            continue
        m = re.fullmatch(r"let (?:mut )?data = get_solid\(\)", s)
        if m:
            uses_get_solid = True
            k = statics["__K"]
            inserts_seq.append(statics["REFE"])
            continue
        m = re.fullmatch(
            r"for kmer in cocktail::tokenizer::Tokenizer::new\(&?(\w+), (\d+)\) \{ data\.set\(kmer, true\)",
            s,
        )
        if m:
            assert int(m.group(2)) == k, (module, name, s)
            inserts_seq.append(env[m.group(1)])
            # the closing brace is glued to the next statement
            if idx < len(stmts) and stmts[idx].startswith("}"):
                stmts[idx] = stmts[idx][1:].strip()
                if not stmts[idx]:
                    idx += 1
            continue
        m = re.fullmatch(r"data\.set\(cocktail::kmer::seq2bit\(" + BYTES + r"\), true\)", s)
        if m:
            assert len(m.group(1)) == k, (module, name, s)
            inserts_kmer.append(m.group(1))
            continue
        if re.fullmatch(r"let set: set::BoxKmerSet = Box::new\(set::Pcon::new\(data\)\)", s):
            continue
        m = re.fullmatch(r"let corrector = (\w+)::new\(&set(?:, (\d+))?(?:, (\d+))?\)", s)
        if m:
            corrector = {"method": m.group(1)}
            args = [int(x) for x in m.groups()[1:] if x is not None]
            if m.group(1) in ("One", "Two", "GapSize"):
                assert len(args) == 1
                corrector["confirm"] = args[0]
            elif m.group(1) == "Greedy":
                assert len(args) == 2
                corrector["max_search"], corrector["nb_validate"] = args
            else:
                assert m.group(1) == "Graph" and not args
            continue
        m = re.fullmatch(
            r"assert_eq!\(&?(\w+), corrector\.correct\(&?(?:(\w+)|filter\((\w+)\))\)\.as_slice\(\)\)", s
        )
        if m:
            exp = env[m.group(1)]
            if m.group(2):
                inp = env[m.group(2)]
            else:
                inp = env[m.group(3)].replace("-", "")
            asserts.append({"input": inp, "expected": exp.replace("-", "")})
            continue
        if s.startswith("println!("):
            continue
        raise SystemExit(f"unparsed statement in {module}::{name}: {s!r}")

    assert k is not None and corrector is not None and asserts, (module, name)
    return {
        "module": module,
        "name": name,
        "k": k,
        "insert_all_kmers_of": inserts_seq,
        "insert_kmers": inserts_kmer,
        "corrector": corrector,
        "asserts": asserts,
    }


def main():
    out = {"generated_by": "tests/golden/make_kats.py", "correctors": [], "helpers": [], "set": []}
    for module, rel in FILES:
        src = (REF / rel).read_text()
        statics = {}
        m = re.search(r"static REFE: &\[u8\] = " + BYTES, src)
        if m:
            statics["REFE"] = m.group(1)
            statics["__K"] = int(re.search(r"static K: u8 = (\d+)", src).group(1))
        for name, ignored, line, body in split_tests(src):
            t = parse_test(module, name, body, statics)
            t["ignored_upstream"] = ignored
            t["source"] = f"src/{rel}:{line}"
            out["correctors"].append(t)

    # src/correct/mod.rs:170-181 — alt_nucs KAT (literal)
    src = (REF / "correct/mod.rs").read_text()
    assert 'seq2bit(b"ACTGA")' in src and 'seq2bit(b"ACTGT")' in src and "vec![0, 2]" in src
    out["helpers"].append(
        {
            "name": "found_alt_kmer",
            "source": "src/correct/mod.rs:170",
            "k": 5,
            "insert_kmers": ["ACTGA", "ACTGT"],
            "alt_nucs_of": "ACTGC",
            "expected": [0, 2],
        }
    )

    # src/set/pcon.rs:198-255 — set KATs: canonical / forward / absence / k
    src = (REF / "set/pcon.rs").read_text()
    m = re.search(r"static SEQ: &\[u8\] = " + BYTES, src)
    out["set"].append(
        {
            "source": "src/set/pcon.rs:202",
            "k": 11,
            "seq": m.group(1),
            "checks": ["canonical_kmers_present", "forward_kmers_present", "get(0)==false", "k()==11"],
        }
    )

    # src/set/hash.rs:185-242 — the same four KATs through Hash::from_fasta (FILE = ">1\n" + sequence)
    src = (REF / "set/hash.rs").read_text()
    m = re.search(r'static FILE: &\[u8\] = b"([^"]*)"', src)
    fasta = m.group(1).encode().decode("unicode_escape")
    assert fasta.startswith(">1\n")
    out["hash_set"] = [
        {
            "source": "src/set/hash.rs:189",
            "k": 11,
            "seq": fasta.split("\n", 1)[1].replace("\n", ""),
            "checks": ["canonical_kmers_present", "forward_kmers_present", "get(0)==false", "k()==11"],
        }
    ]

    n = len(out["correctors"])
    per = {}
    for t in out["correctors"]:
        per[t["module"]] = per.get(t["module"], 0) + 1
    print(f"{n} corrector KATs: {per}", file=sys.stderr)
    dst = Path(__file__).with_name("kats.json")
    dst.write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
