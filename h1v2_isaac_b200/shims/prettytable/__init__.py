"""Minimal prettytable stand-in (reference use: biped_tasks/utils/cat/constraint_manager.py:12, manager __str__ tables)."""


class PrettyTable:
    def __init__(self, field_names=None):
        self.title, self.field_names, self.align, self._rows = "", list(field_names or []), {}, []

    def add_row(self, row):
        self._rows.append([str(x) for x in row])

    def get_string(self):
        rows = [[str(f) for f in self.field_names]] + self._rows
        w = [max(len(r[i]) for r in rows) for i in range(len(self.field_names))] if self.field_names else []
        out = [self.title] if self.title else []
        out += [" | ".join(c.ljust(w[i]) for i, c in enumerate(r)) for r in rows]
        return "\n".join(out)

    __str__ = get_string
