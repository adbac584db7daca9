"""Minimal .xlsx reader/writer on the Python standard library.

The reference loads grids with ``pandas.read_excel`` (classes.py:126-128) and
writes results with ``pandas.ExcelWriter`` (helm.py:816-848).  Neither openpyxl
nor xlrd is available in this image, so the boundary is served by a small
zipfile + ElementTree implementation that understands exactly what the
reference's files contain: numeric ``<v>`` cells, ``t="inlineStr"`` cells and
``t="s"`` shared strings; sheets resolved by name through ``xl/workbook.xml``
and ``xl/_rels/workbook.xml.rels``.
"""

import re
import zipfile
import xml.etree.ElementTree as ET
from xml.sax.saxutils import escape

import numpy as np

_NS = "{http://schemas.openxmlformats.org/spreadsheetml/2006/main}"
_RNS = "{http://schemas.openxmlformats.org/officeDocument/2006/relationships}"
_PNS = "{http://schemas.openxmlformats.org/package/2006/relationships}"
_CELL = re.compile(r"([A-Z]+)(\d+)")


def _col_index(letters):
    c = 0
    for ch in letters:
        c = c * 26 + (ord(ch) - 64)
    return c - 1


def _col_letters(idx):
    s = ""
    idx += 1
    while idx:
        idx, r = divmod(idx - 1, 26)
        s = chr(65 + r) + s
    return s


def sheet_names(path):
    with zipfile.ZipFile(path) as z:
        wb = ET.fromstring(z.read("xl/workbook.xml"))
    return [s.get("name") for s in wb.iter(_NS + "sheet")]


def read_sheet(path, sheet_name):
    """Return the sheet as a list of rows; each row a list of float | str | None."""
    with zipfile.ZipFile(path) as z:
        wb = ET.fromstring(z.read("xl/workbook.xml"))
        rid = None
        for s in wb.iter(_NS + "sheet"):
            if s.get("name") == sheet_name:
                rid = s.get(_RNS + "id")
                break
        if rid is None:
            raise KeyError("sheet %r not found in %s" % (sheet_name, path))
        rels = ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))
        target = None
        for r in rels.iter(_PNS + "Relationship"):
            if r.get("Id") == rid:
                target = r.get("Target")
                break
        target = target.lstrip("/")
        if not target.startswith("xl/"):
            target = "xl/" + target
        shared = None
        if "xl/sharedStrings.xml" in z.namelist():
            sst = ET.fromstring(z.read("xl/sharedStrings.xml"))
            shared = ["".join(t.text or "" for t in si.iter(_NS + "t"))
                      for si in sst.iter(_NS + "si")]
        root = ET.fromstring(z.read(target))
    rows = []
    for row in root.iter(_NS + "row"):
        r_idx = int(row.get("r")) - 1
        while len(rows) <= r_idx:
            rows.append([])
        cur = rows[r_idx]
        for c in row.iter(_NS + "c"):
            m = _CELL.match(c.get("r"))
            ci = _col_index(m.group(1))
            while len(cur) <= ci:
                cur.append(None)
            t = c.get("t", "n")
            if t == "inlineStr":
                cur[ci] = "".join(x.text or "" for x in c.iter(_NS + "t"))
            else:
                v = c.find(_NS + "v")
                if v is None or v.text is None:
                    continue
                if t == "s":
                    cur[ci] = shared[int(v.text)]
                elif t in ("str", "e"):
                    cur[ci] = v.text
                elif t == "b":
                    cur[ci] = float(int(v.text))
                else:
                    cur[ci] = float(v.text)
    return rows


def read_numeric_sheet(path, sheet_name):
    """Header-less all-numeric sheet -> float64 ndarray (rows, cols); blanks = 0."""
    rows = [r for r in read_sheet(path, sheet_name) if len(r)]
    ncol = max((len(r) for r in rows), default=0)
    out = np.zeros((len(rows), ncol), dtype=np.float64)
    for i, r in enumerate(rows):
        for j, v in enumerate(r):
            if v is not None:
                out[i, j] = float(v)
    return out


def read_table_with_header(path, sheet_name):
    """First row = header, first column = row index (as pandas writes it).

    Returns dict header -> list of cell values (float or str).
    """
    rows = [r for r in read_sheet(path, sheet_name)]
    header = rows[0]
    cols = {}
    for ci, name in enumerate(header):
        if name is None:
            continue
        cols[name] = [(r[ci] if ci < len(r) else None) for r in rows[1:] if len(r)]
    return cols


def write_xlsx(path, sheets):
    """Write ``sheets`` = [(name, header list, row-index list, columns list-of-lists)].

    Layout equals what ``DataFrame.to_excel`` produces for the reference's
    result files (helm.py:816-841): header row in B1.., index in column A,
    strings as inline strings, numbers as ``t="n"``.
    """
    def cell(ref, v):
        if v is None:
            return ""
        if isinstance(v, str):
            return '<c r="%s" t="inlineStr"><is><t>%s</t></is></c>' % (ref, escape(v))
        return '<c r="%s" t="n"><v>%s</v></c>' % (ref, repr(float(v)))

    content_types = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
                     '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">',
                     '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>',
                     '<Default Extension="xml" ContentType="application/xml"/>',
                     '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/>']
    wb = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
          '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
          'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets>']
    rels = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
            '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">']
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        for k, (name, header, index, columns) in enumerate(sheets, start=1):
            content_types.append(
                '<Override PartName="/xl/worksheets/sheet%d.xml" ContentType="application/'
                'vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>' % k)
            wb.append('<sheet name="%s" sheetId="%d" r:id="rId%d"/>' % (escape(name), k, k))
            rels.append('<Relationship Id="rId%d" Type="http://schemas.openxmlformats.org/'
                        'officeDocument/2006/relationships/worksheet" Target="worksheets/sheet%d.xml"/>' % (k, k))
            body = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
                    '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>']
            body.append('<row r="1">' + "".join(
                cell("%s1" % _col_letters(j + 1), h) for j, h in enumerate(header)) + '</row>')
            for i, idx in enumerate(index):
                r = i + 2
                cells = [cell("A%d" % r, idx)]
                for j, col in enumerate(columns):
                    cells.append(cell("%s%d" % (_col_letters(j + 1), r), col[i]))
                body.append('<row r="%d">%s</row>' % (r, "".join(cells)))
            body.append('</sheetData></worksheet>')
            z.writestr("xl/worksheets/sheet%d.xml" % k, "".join(body))
        content_types.append('</Types>')
        wb.append('</sheets></workbook>')
        rels.append('</Relationships>')
        z.writestr("[Content_Types].xml", "".join(content_types))
        z.writestr("_rels/.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/'
                   'relationships/officeDocument" Target="xl/workbook.xml"/></Relationships>')
        z.writestr("xl/workbook.xml", "".join(wb))
        z.writestr("xl/_rels/workbook.xml.rels", "".join(rels))


def write_numeric_xlsx(path, sheets):
    """Header-less numeric sheets (the reference's *case* layout).

    ``sheets`` = [(name, 2-D array)].  Cells start at A1.
    """
    wrapped = []
    content = []
    for name, arr in sheets:
        content.append((name, np.asarray(arr, dtype=np.float64)))
    content_types = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
                     '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">',
                     '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>',
                     '<Default Extension="xml" ContentType="application/xml"/>',
                     '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/>']
    wb = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
          '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
          'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets>']
    rels = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
            '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">']
    del wrapped
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        for k, (name, arr) in enumerate(content, start=1):
            content_types.append(
                '<Override PartName="/xl/worksheets/sheet%d.xml" ContentType="application/'
                'vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>' % k)
            wb.append('<sheet name="%s" sheetId="%d" r:id="rId%d"/>' % (escape(name), k, k))
            rels.append('<Relationship Id="rId%d" Type="http://schemas.openxmlformats.org/'
                        'officeDocument/2006/relationships/worksheet" Target="worksheets/sheet%d.xml"/>' % (k, k))
            body = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
                    '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>']
            for i in range(arr.shape[0]):
                r = i + 1
                cells = []
                for j in range(arr.shape[1]):
                    v = arr[i, j]
                    txt = repr(int(v)) if float(v).is_integer() and abs(v) < 1e15 else repr(float(v))
                    cells.append('<c r="%s%d"><v>%s</v></c>' % (_col_letters(j), r, txt))
                body.append('<row r="%d">%s</row>' % (r, "".join(cells)))
            body.append('</sheetData></worksheet>')
            z.writestr("xl/worksheets/sheet%d.xml" % k, "".join(body))
        content_types.append('</Types>')
        wb.append('</sheets></workbook>')
        rels.append('</Relationships>')
        z.writestr("[Content_Types].xml", "".join(content_types))
        z.writestr("_rels/.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/'
                   'relationships/officeDocument" Target="xl/workbook.xml"/></Relationships>')
        z.writestr("xl/workbook.xml", "".join(wb))
        z.writestr("xl/_rels/workbook.xml.rels", "".join(rels))
