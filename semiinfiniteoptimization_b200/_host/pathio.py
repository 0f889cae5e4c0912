"""Wire formats of the reference's trajectory inputs and outputs:

    *.path     Klampt linear path, one milestone per line:   "<time>\\t<n> <q1> ... <qn>"
               (what `Trajectory.load / save` of klampt.model.trajectory read and write; the reference loads
               resources/robottrajopt_initial.path at robottrajopt.py:16-17, 71 and saves output/*.path)
    *.configs  Klampt `Configs` resource, one configuration per line:   "<n>\\t<q1> ... <qn>"
               (resources/robotplanopt_chair_start_goal.configs, read by robotplanopt.py:67-80)

Times are written with "%f"; values the way the reference's interpreter (Python 2) printed a float -- `str(x)`, i.e. 12
significant digits with a trailing ".0" on integral values -- so that a shipped file read and written back is the same
file byte for byte (tests/test_planopt_cpu.py).  `precise=True` writes `repr(x)` instead (round-trips every double).
"""


def format_value(x, precise=False):
    x = float(x)
    if precise:
        return repr(x)
    s = "%.12g" % x
    if "." not in s and "e" not in s and "n" not in s:          # 'inf' / 'nan' contain an n
        s += ".0"
    return s


def path_to_text(times, milestones, precise=False):
    return "".join("%f\t%d %s\n" % (t, len(m), " ".join(format_value(v, precise) for v in m)) for t, m in zip(times, milestones))


def path_from_text(text):
    times, ms = [], []
    for ln, line in enumerate(text.splitlines(), 1):
        tok = line.split()
        if not tok:
            continue
        if len(tok) < 2:
            raise ValueError("line %d: expected '<time> <n> <values...>'" % ln)
        n = int(tok[1])
        if len(tok) != 2 + n:
            raise ValueError("line %d: %d values announced, %d present" % (ln, n, len(tok) - 2))
        times.append(float(tok[0]))
        ms.append([float(v) for v in tok[2:]])
    if any(b < a for a, b in zip(times, times[1:])):
        raise ValueError("milestone times must be non-decreasing")
    return times, ms


def configs_to_text(configs, precise=False):
    return "".join("%d\t%s\n" % (len(q), " ".join(format_value(v, precise) for v in q)) for q in configs)


def configs_from_text(text):
    out = []
    for ln, line in enumerate(text.splitlines(), 1):
        tok = line.split()
        if not tok:
            continue
        n = int(tok[0])
        if len(tok) != 1 + n:
            raise ValueError("line %d: %d values announced, %d present" % (ln, n, len(tok) - 1))
        out.append([float(v) for v in tok[1:]])
    return out


def load_path(fn):
    with open(fn) as f:
        return path_from_text(f.read())


def save_path(fn, times, milestones, precise=False):
    with open(fn, "w") as f:
        f.write(path_to_text(times, milestones, precise))


def load_configs(fn):
    with open(fn) as f:
        return configs_from_text(f.read())


def save_configs(fn, configs, precise=False):
    with open(fn, "w") as f:
        f.write(configs_to_text(configs, precise))
