"""Text output of sample matrices on the device.

``savetxt(fname, X, delimiter=" ", precision=8)`` writes exactly the bytes of
``np.savetxt(fname, X, delimiter=delimiter, fmt="%.<precision>e")`` -- the call the reference's CLI makes
(sample/sobol.py:247-251 and the other ``cli_action``s) -- but formats on the GPU: ``sa_format_values_f64``
(exact "%.Pe" of every value), a prefix sum of the row lengths, ``sa_pack_text_rows``.  Rows are processed in
chunks so that the 32-byte slots of a chunk stay bounded."""
import numpy as np
import torch

from .. import _lib, _runtime as rt

__all__ = ["savetxt", "format_rows", "loadtxt", "parse_text"]

CHUNK_VALUES = 1 << 24     # values formatted per pass (0.5 GB of slots)


def format_rows(X, delimiter=" ", precision=8):
    """Device text of a [rows, cols] block: returns a uint8 CUDA tensor."""
    if len(delimiter.encode()) != 1:
        raise ValueError("delimiter must be a single byte")
    Xd = rt.to_device(X, torch.float64)
    if Xd.ndim == 1:
        Xd = Xd.reshape(-1, 1)      # np.savetxt writes a 1-D array one value per line
    if Xd.ndim != 2:
        raise ValueError(f"Expected 1D or 2D array, got {Xd.ndim}D array instead")
    rows, cols = int(Xd.shape[0]), int(Xd.shape[1])
    if rows == 0 or cols == 0:
        return rt.empty((0,), torch.uint8)
    slot = int(_lib.lib().sa_text_slot_bytes())
    slots = rt.empty((rows * cols * slot,), torch.uint8)
    lens = rt.empty((rows * cols,), torch.uint8)
    rowlen = rt.empty((rows,), torch.int64)
    _lib.call("sa_format_values_f64", rt.stream_ptr(), rt.ptr(Xd), rows, cols, int(precision), rt.ptr(slots),
              rt.ptr(lens), rt.ptr(rowlen))
    rowend = torch.cumsum(rowlen, dim=0)
    rowoff = (rowend - rowlen).contiguous()
    text = rt.empty((int(rowend[-1].item()),), torch.uint8)
    _lib.call("sa_pack_text_rows", rt.stream_ptr(), rt.ptr(slots), rt.ptr(lens), rows, cols,
              ord(delimiter), rt.ptr(rowoff), rt.ptr(text))
    return text


def savetxt(fname, X, delimiter=" ", precision=8):
    """``np.savetxt(fname, X, delimiter=delimiter, fmt="%.{precision}e")``, formatted on the device.
    ``fname``: path or binary file object.  X: host array or CUDA tensor, 1-D or 2-D."""
    rt.require_cuda()
    own = isinstance(fname, (str, bytes)) or hasattr(fname, "__fspath__")
    fh = open(fname, "wb") if own else fname
    try:
        Xv = X if isinstance(X, torch.Tensor) else np.asarray(X, dtype=np.float64)
        if Xv.ndim == 0:
            Xv = Xv.reshape(1, 1)
        rows = int(Xv.shape[0])
        cols = int(Xv.shape[1]) if Xv.ndim == 2 else 1
        step = max(1, CHUNK_VALUES // max(1, cols))
        for r0 in range(0, rows, step):
            text = format_rows(Xv[r0:r0 + step], delimiter, precision)
            fh.write(text.cpu().numpy().tobytes())
    finally:
        if own:
            fh.close()


def parse_text(data, delimiter=None):
    """bytes / uint8 array of numeric text -> CUDA float64 tensor [rows, cols] (``sa_text_*`` kernels)."""
    rt.require_cuda()
    if delimiter is not None and len(delimiter.encode()) != 1:
        raise ValueError("delimiter must be a single byte")
    delim = 0 if delimiter in (None, " ", "\t") else ord(delimiter)
    if delim in (10, 35):
        raise ValueError("delimiter cannot be a newline or '#'")
    if isinstance(data, torch.Tensor):
        text = rt.to_device(data, torch.uint8).reshape(-1)
    else:
        text = rt.to_device(np.frombuffer(data, dtype=np.uint8), torch.uint8)
    n = int(text.numel())
    if n == 0:
        return rt.empty((0, 0), torch.float64)
    L = _lib.lib()
    chunk = int(L.sa_text_parse_chunk_bytes())
    blockcount = rt.empty((-(-n // chunk),), torch.int64)
    _lib.call("sa_text_count_newlines", rt.stream_ptr(), rt.ptr(text), n, rt.ptr(blockcount))
    blockend = torch.cumsum(blockcount, dim=0)
    nlines = int(blockend[-1].item()) + 1
    blockoff = (blockend - blockcount).contiguous()
    linestart = rt.empty((nlines,), torch.int64)
    _lib.call("sa_text_line_starts", rt.stream_ptr(), rt.ptr(text), n, rt.ptr(blockoff), rt.ptr(linestart))
    ntok = rt.empty((nlines,), torch.int32)
    flag = rt.zeros((1,), torch.int32)
    _lib.call("sa_text_parse_lines", rt.stream_ptr(), rt.ptr(text), n, rt.ptr(linestart), nlines, delim, 0,
              rt.ptr(ntok), None, 0, None, rt.ptr(flag))
    nonblank = ntok > 0
    rows = int(nonblank.sum().item())
    if rows == 0:
        return rt.empty((0, 0), torch.float64)
    cols = int(ntok.max().item())
    rowidx = (torch.cumsum(nonblank.to(torch.int64), dim=0) - 1).contiguous()
    out = rt.empty((rows, cols), torch.float64)
    _lib.call("sa_text_parse_lines", rt.stream_ptr(), rt.ptr(text), n, rt.ptr(linestart), nlines, delim, 1,
              None, rt.ptr(rowidx), cols, rt.ptr(out), rt.ptr(flag))
    f = int(flag.item())
    if f & 2:
        raise ValueError("the number of columns changed between rows")
    if f & 1:
        raise ValueError("could not convert a field to float")
    if f & 4:
        raise ValueError("a field has more than 40 significant digits")
    return out


def loadtxt(fname, delimiter=None, usecols=None, ndmin=0, device=False):
    """``np.loadtxt(fname, delimiter=delimiter, usecols=usecols, ndmin=ndmin)`` for numeric text, parsed on
    the device.  Returns an ndarray (or the CUDA tensor with ``device=True``), squeezed like NumPy's."""
    if hasattr(fname, "read"):
        data = fname.read()
        if isinstance(data, str):
            data = data.encode()
    else:
        with open(fname, "rb") as fh:
            data = fh.read()
    out = parse_text(data, delimiter)
    if usecols is not None:
        cols = [usecols] if isinstance(usecols, int) else list(usecols)
        out = out[:, cols]
    if device:
        return out
    arr = out.cpu().numpy()
    if arr.ndim > ndmin:
        arr = np.squeeze(arr)
    while arr.ndim < ndmin:
        arr = arr[None] if arr.ndim else arr.reshape(1)
    return arr
