"""Console / log.txt reporting with the surface of the reference's ``ScreenPrinter`` (safeincave/ScreenOutput.py:38-571).

What user tooling relies on, and what is therefore reproduced exactly:

* the run prints a framed report -- banner, ``Mesh info``, ``Partition(s) info``, ``Solver info``, ``Constitutive model``,
  ``Output info`` -- followed by one table row per time step ``| step | dt | t / t_final | # of iters | error |`` and a
  closing ``Total time: HH:MM:SS (s seconds)`` line (ScreenOutput.py:105-132, 355-377);
* everything printed is also kept in ``self.log`` and written to ``<output_folder>/log.txt`` of every ``SaveFields``
  object on ``close()`` (:379-393).  Post-processing scripts parse that file: the Newton iterations per step are the 4th
  ``|``-separated cell of every row whose first cell is an integer
  (examples/mechanics/nobian/Simulation/Run_sensitivity.py:353-373), and the mesh sizes are read from the 4th line after
  ``| Mesh info:`` (examples/mechanics/4_cavern/plot_results.py:139-150);
* the frame is 97 characters wide; a table narrower than the frame is closed with `` |`` and padded to the frame
  (:455-506), cells are rendered with printf-style formats and left / center / right alignment (:538-571).

Only rank 0 prints.  Partition sizes come from the cell partition of ``safeincave_b200.partition`` instead of DOLFINx's
index maps.  This is host-side reporting: nothing here touches the device.
"""
from __future__ import annotations

import os
import sys
import time

FRAME_WIDTH = 97


class _Table:
    """Column layout of the table currently being printed: widths are those of the header cells."""

    def __init__(self, header, align):
        self.header = [str(h) for h in header]
        self.widths = [len(h) for h in self.header]
        self.header_align = [align] * len(self.header)

    @staticmethod
    def cell(value, width, align, fmt=None):
        text = (fmt % value) if fmt is not None else value
        spec = {"left": "<", "center": "^"}.get(align, ">")
        return format(text, f"{spec}{width}")

    def frame(self, line):
        """Close a table line and pad it to the frame."""
        room = FRAME_WIDTH - len(line) - 1
        if room > 1:
            line += " |"
            room = FRAME_WIDTH - len(line) - 1
        return line + " " * room + "|"

    def divider(self):
        line = "+" + "+".join("-" * (w + 2) for w in self.widths) + "+"
        room = FRAME_WIDTH - len(line) - 1
        if room > -1:
            return line + "-" * room + "+"
        return line + "-" * (room - 1)

    def line(self, values, aligns, formats=None):
        formats = formats if formats is not None else [None] * len(values)
        cells = [self.cell(v, w, a, f) for v, w, a, f in zip(values, self.widths, aligns, formats)]
        return self.frame("| " + " | ".join(cells))


class ScreenPrinter:
    """``ScreenPrinter(grid, solver, material, outputs, time_unit)``: constructing it prints the report header and opens
    the time-step table; ``print_row([...])`` adds a step; ``close()`` prints the total time and writes ``log.txt``."""

    _instance = None

    @classmethod
    def reset_instance(cls):
        """The reference wraps the class in a singleton and resets it per simulator (Simulators.py:89, 307)."""
        cls._instance = None

    def __init__(self, grid, solver, material, outputs, time_unit: str = "hour", stream=None):
        self.master_division_plus = "+" + "-" * (FRAME_WIDTH - 2) + "+"
        self.master_division = "-" * (FRAME_WIDTH - 2)
        self.max_width = FRAME_WIDTH
        self.log = ""
        self.grid, self.solver, self.mat = grid, solver, material
        self.outputs = list(outputs) if outputs is not None else []
        self.time_unit = time_unit
        self.stream = stream if stream is not None else sys.stdout
        self.output_folders = []
        self.row_formats, self.row_align = [], []
        self._table = _Table(["", ""], "left")
        ScreenPrinter._instance = self
        self.set_welcome()
        self.print_welcome()
        self.print_mesh_info()
        self.print_partition_info()
        self.print_solver_info()
        self.print_constitutive_model()
        self.print_output_info()
        self.begin()

    # ------------------------------------------------------------------ plumbing
    @property
    def _rank(self):
        try:
            return int(self.grid.mesh.comm.rank)
        except AttributeError:
            return 0

    @property
    def divider(self):
        return self._table.divider()

    @property
    def widths(self):
        return self._table.widths

    @property
    def header_columns(self):
        return self._table.header

    def add_to_log(self, message: str) -> None:
        self.log += "\n" + message

    def print_on_screen(self, raw_comment: str) -> None:
        if self._rank != 0:
            return
        print(raw_comment, file=self.stream)
        self.stream.flush()
        self.add_to_log(raw_comment)

    def print_comment(self, comment, align: str = "left") -> None:
        if comment is not None:
            self.print_on_screen("|" + _Table.cell(comment, FRAME_WIDTH - 2, align) + "|")

    def format_cell(self, text, width: int, alignment: str, text_format: str = None):
        return _Table.cell(text, width, alignment, text_format)

    def make_divider(self, widths, middle: str = "+"):
        t = _Table([" " * w for w in widths], "left")
        return t.divider() if middle == "+" else t.divider().replace("+", middle)

    def set_header_columns(self, header_columns, align: str) -> None:
        self._table = _Table(header_columns, align)

    def set_row_formats(self, row_formats, row_align) -> None:
        self.row_formats, self.row_align = list(row_formats), list(row_align)

    def print_header(self) -> None:
        t = self._table
        self.print_on_screen(t.divider())
        self.print_on_screen(t.line(t.header, t.header_align))
        self.print_on_screen(t.divider())

    def print_row(self, values) -> None:
        self.print_on_screen(self._table.line(values, self.row_align, self.row_formats))

    def _section(self, title, header, header_align, formats, aligns, rows):
        self.print_comment(title)
        self.set_header_columns(header, header_align)
        self.print_header()
        self.set_row_formats(formats, aligns)
        for row in rows:
            self.print_row(row)
        self.print_on_screen(self.divider)
        self.print_comment(" ")

    # ------------------------------------------------------------------ report sections
    def set_welcome(self) -> None:
        bar = "+" + "=" * (FRAME_WIDTH - 2) + "+"
        title = "S A F E   I N   C A V E   --   B200-native mechanics path (safeincave_b200)"
        self.welcome_text = "\n".join([bar, "|" + " " * (FRAME_WIDTH - 2) + "|",
                                       "|" + _Table.cell(title, FRAME_WIDTH - 2, "center") + "|",
                                       "|" + " " * (FRAME_WIDTH - 2) + "|", bar])

    def print_welcome(self) -> None:
        self.print_on_screen(self.welcome_text)
        self.print_comment(" ")

    def _global_sizes(self):
        part = getattr(self.grid, "partition", None)
        if part is not None:
            return int(part.n_global_cells), int(part.n_global_nodes)
        tm = self.grid.tetmesh
        return int(tm.n_cells), int(tm.n_nodes)

    def print_mesh_info(self) -> None:
        folder = str(getattr(self.grid, "grid_folder", "") or "(in memory)")
        pad = max(len(folder) - len("Location"), 0)
        n_elems, n_nodes = self._global_sizes()
        self._section(" Mesh info:", ["# of elements", "# of nodes", "Location" + " " * pad], "left",
                      ["%.i", "%.i", "%s"], ["left"] * 3, [[n_elems, n_nodes, folder]])

    def print_partition_info(self) -> None:
        part = getattr(self.grid, "partition", None)
        rows = []
        if part is None:
            tm = self.grid.tetmesh
            rows.append([1, int(tm.n_cells), int(tm.n_nodes)])
        else:       # contiguous Morton-curve chunks of the cells, interface nodes duplicated (partition.py): every rank
            from .partition import chunk_bounds          # knows the whole plan, nothing is gathered
            b = chunk_bounds(int(part.n_global_cells), int(part.n_ranks))
            for r in range(int(part.n_ranks)):
                n_nodes = int(((part.touch >> r) & 1).sum()) if part.touch is not None else 0
                rows.append([r + 1, b[r + 1] - b[r], n_nodes])
        self._section(" Partition(s) info:", ["Partition #", "# of elements", "# of nodes"], "left",
                      ["%.i", "%.i", "%.i"], ["center"] * 3, rows)

    def print_solver_info(self) -> None:
        rtol, _atol, _divtol, max_it = self.solver.getTolerances()
        self._section(" Solver info:", ["KSP_type", "PC_type", "  rtol  ", "max_it"], "center",
                      ["%s", "%s", "%.1e", "%.i"], ["center"] * 4,
                      [[self.solver.getType(), self.solver.getPC().getType(), rtol, max_it]])

    def print_constitutive_model(self) -> None:
        groups = [("elastic", getattr(self.mat, "elems_e", [])), ("non-elastic", getattr(self.mat, "elems_ne", [])),
                  ("thermoelastic", getattr(self.mat, "elems_th", []))]
        if sum(len(g) for _, g in groups) == 0:
            return
        names = [", ".join(str(e.name) for e in g) for _, g in groups]
        pad = max(max(len(n) for n in names) - len("List of elements"), 0)
        self._section(" Constitutive model:", ["Element type ", "List of elements" + " " * pad], "left",
                      ["%s", "%s"], ["left"] * 2, [[kind, n] for (kind, _), n in zip(groups, names)])

    def print_output_info(self) -> None:
        rows = []
        self.output_folders = []
        for output in self.outputs:
            self.output_folders.append(output.output_folder)
            for fd in output.fields_data:
                rows.append([output.output_folder, fd["field_name"], fd["label_name"]])
        self._section(" Output info:", ["Location" + 10 * " ", "Field name      ", "Label name             "], "center",
                      ["%s", "%s", "%s"], ["left"] * 3, rows)

    # ------------------------------------------------------------------ the time-step table
    def begin(self) -> None:
        self.start_timer()
        self.set_header_columns(["Step counter", f"dt ({self.time_unit})", f"t / t_final ({self.time_unit})",
                                 "# of iters", "Non-linear error"], "center")
        self.set_row_formats(["%i", "%.3f", "%s", "%.i", "%.4e"], ["center"] * 5)
        self.print_header()

    def start_timer(self) -> None:
        self.start = time.perf_counter()

    def close(self) -> None:
        self.print_on_screen(self.divider)
        if self._rank != 0:
            return
        self.final = time.perf_counter()
        cpu_time = self.final - self.start
        stamp = time.strftime("%H:%M:%S", time.gmtime(cpu_time))
        self.print_on_screen(_Table.cell(f"Total time: {stamp} ({cpu_time} seconds)", FRAME_WIDTH, "right"))
        for folder in self.output_folders:
            self.save_log(folder)

    def save_log(self, output_folder: str) -> None:
        os.makedirs(output_folder, exist_ok=True)
        with open(os.path.join(output_folder, "log.txt"), "w") as fh:
            fh.write(self.log)
