"""oracle shim"""
def cprint(*a, **k): print(*a[:1])
def colored(s, *a, **k): return s
