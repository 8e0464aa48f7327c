"""TEST INFRASTRUCTURE: CPU restatement of the output formatter row, the checker of cco_format_es_bulk.

What it restates, per primary item (row) of the indicator model:
  IndexedDatasetConversions.toStringMapRDD  /root/reference/src/main/scala/package.scala:82-110
      non-zeros ordered by -LLR (the model rows already are: llr desc, col asc), mapped to column id strings, LLR dropped;
      a row without cells gives an empty array
  URModel.save: one map per item, all event fields merged, plus "id" -> item   /root/reference/src/main/scala/URModel.scala:47-102
  EsClient: saveToEs with "es.mapping.id" -> "id" (bulk index action per document)   /root/reference/src/main/scala/EsClient.scala:300-313
Bytes per document (what elasticsearch-hadoop puts on the wire for an index action with a mapped id):
    {"index":{"_id":"<item>"}}\n{"id":"<item>","<event 0>":["<col>",...],"<event 1>":[...]}\n
String escaping: '"' and '\\' get a backslash, bytes below 0x20 become \\u00xx (lower-case hex), everything else passes
through as UTF-8.  Parity is pinned on the structure (json.loads of every line equals the host mirror's to_string_map on the
reference fixtures, tests/test_format.py); the byte layout is this repo's definition, identical in the CUDA path.
"""
from __future__ import annotations


def json_escape(s: str) -> bytes:
    out = bytearray()
    for ch in s.encode("utf-8"):
        if ch in (0x22, 0x5C):
            out += b"\\" + bytes([ch])
        elif ch < 0x20:
            out += b"\\u00" + b"0123456789abcdef"[ch >> 4:(ch >> 4) + 1] + b"0123456789abcdef"[ch & 15:(ch & 15) + 1]
        else:
            out.append(ch)
    return bytes(out)


def es_bulk(indicators, names, row_ids, col_ids, row_begin: int = 0, row_end: int | None = None) -> bytes:
    """indicators: per event (row_ptr, col_idx) of rows [row_begin, row_end) (row_ptr relative to row_begin);
    names: event names; row_ids: id strings of the primary item space; col_ids[i]: id strings of event i's item space."""
    n_rows = len(indicators[0][0]) - 1
    if row_end is None:
        row_end = row_begin + n_rows
    assert row_end - row_begin == n_rows
    esc_rows = [json_escape(x) for x in row_ids]
    esc_cols = [[json_escape(x) for x in ids] for ids in col_ids]
    esc_names = [json_escape(x) for x in names]
    out = bytearray()
    for r in range(n_rows):
        rid = esc_rows[row_begin + r]
        out += b'{"index":{"_id":"' + rid + b'"}}\n{"id":"' + rid + b'"'
        for i, (rp, ci) in enumerate(indicators):
            out += b',"' + esc_names[i] + b'":['
            out += b",".join(b'"' + esc_cols[i][int(c)] + b'"' for c in ci[int(rp[r]) - int(rp[0]):int(rp[r + 1]) - int(rp[0])])
            out += b"]"
        out += b"}\n"
    return bytes(out)
