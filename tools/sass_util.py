"""nvdisasm --print-line-info output of a cubin with several kernels: pick the lines of the function an ncu source page covers."""
import re

INS_RE = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
SEC_RE = re.compile(r'^\s*\.text\.(\S+):')


def function_lines(sass_path: str, n_rows: int):
    """Lines of the .text section whose instruction count equals n_rows (the ncu page's row count); if no section matches
    (or the file has a single function) all lines are returned."""
    secs, cur = {}, None
    for ln in open(sass_path):
        m = SEC_RE.match(ln)
        if m:
            cur = m.group(1); secs[cur] = []
        if cur is not None:
            secs[cur].append(ln)
    for name, lines in secs.items():
        if sum(1 for l in lines if INS_RE.match(l)) == n_rows:
            return lines
    return [l for lines in secs.values() for l in lines] if secs else list(open(sass_path))
