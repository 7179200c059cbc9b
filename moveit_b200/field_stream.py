"""On-disk format of PropagationDistanceField::writeToStream / readFromStream
(propagation_distance_field.cpp:636-761): seven text lines written with the default ostream formatting of a double
(six significant digits, "%g"), then ONE zlib stream (boost::iostreams::zlib_compressor defaults: zlib header,
default level) of the occupancy bits, x-major, y, then z in groups of eight, bit zi <-> cell z+zi, set iff d^2 == 0.
The bit packing itself runs on the GPU (b200df_field_pack_occupancy / unpack); this module is the host-side framing.
"""
import zlib

_KEYS = ("resolution", "size_x", "size_y", "size_z", "origin_x", "origin_y", "origin_z")


def _fmt(v):
    return "%g" % float(v)  # std::ostream << double with the default precision (6)


def write_stream(fileobj, resolution, size, origin, payload):
    """writeToStream :636-673.  payload: the packed occupancy bytes (numpy uint8 array or bytes)."""
    values = (resolution, size[0], size[1], size[2], origin[0], origin[1], origin[2])
    for key, v in zip(_KEYS, values):
        fileobj.write(("%s: %s\n" % (key, _fmt(v))).encode("ascii"))
    fileobj.write(zlib.compress(bytes(payload)))
    return True


def read_stream(fileobj):
    """readFromStream :675-761.  Returns (header dict, payload bytes) or None when the stream is not a field
    (the reference returns false).  The header values are the six-digit roundings that were written."""
    data = fileobj.read()
    pos = 0
    header = {}
    for key in _KEYS:
        # operator>>(std::string) then operator>>(double): tokens separated by whitespace
        end = data.find(b"\n", pos)
        if end < 0:
            return None
        toks = data[pos:end].split()
        if len(toks) != 2 or toks[0] != (key + ":").encode("ascii"):
            return None
        try:
            header[key] = float(toks[1])
        except ValueError:
            return None
        pos = end + 1  # "this should be newline" (:717-719)
    try:
        payload = zlib.decompress(data[pos:])
    except zlib.error:
        return None
    return header, payload


def payload_size(dims):
    nx, ny, nz = dims
    return nx * ny * ((nz + 7) // 8)
