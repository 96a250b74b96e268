"""Synthetic workloads of the BASELINE.json shapes: graph builders and parameter generators shared by
tests/, bench.py, __graft_entry__.smoke() and tools/.  They drive any object with the renderer interface
(on_add_node / on_add_edge / define_* / fill_buffer) — the B200 renderer or the CPU oracle — and import neither."""
