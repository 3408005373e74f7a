"""Static occupancy map (reference ``envs/Map.py:5-64``): 16 m x 16 m at 0.1 m/cell by default.  The dynamic part
(agent discs) is rasterised on the device every step; this class only holds the static layer and converts it to
the row-major bitmap the library reads (``cav_buffers_t.static_map_bits``)."""
import numpy as np


def load_image_map(map_filename, dims):
    """The reference's image path (Map.py:17-21): ``imageio.imread`` -> ``scipy.misc.imresize(.., dims,
    interp='nearest')`` when the shape differs -> ``np.invert(..).astype(bool)`` (occupied = any pixel that is not pure
    white).  ``scipy.misc.imresize`` (removed from SciPy 1.3) was ``PIL.Image.resize(resample=NEAREST)`` on the 8-bit
    image with size ``(dims[1], dims[0])``; Pillow is used directly here.  The reference's own maps
    (``world_maps/00{0,1,2}.png``) are single-channel 8-bit; colour images are reduced to luminance first (the reference
    would carry the channel axis into the map and fail later)."""
    try:
        from PIL import Image
    except ImportError as e:  # pragma: no cover
        raise ImportError("image maps need Pillow (the reference needs imageio + scipy.misc.imresize, Map.py:17-21); "
                          "pass a boolean array or a .npy file instead") from e
    img = Image.open(map_filename)
    if img.mode != "L":
        img = img.convert("L")
    if (img.size[1], img.size[0]) != tuple(dims):
        img = img.resize((dims[1], dims[0]), resample=Image.NEAREST)
    return np.invert(np.asarray(img, dtype=np.uint8)).astype(bool)


class Map(object):
    def __init__(self, x_width, y_width, grid_cell_size, map_filename=None, static_map=None):
        self.x_width, self.y_width, self.grid_cell_size = x_width, y_width, grid_cell_size
        dims = (int(x_width / grid_cell_size), int(y_width / grid_cell_size))
        if static_map is not None:
            sm = np.asarray(static_map).astype(bool)
            assert sm.shape == dims, "static map must be %s" % (dims,)
            self.static_map = sm
        elif map_filename is None:
            self.static_map = np.zeros(dims, dtype=bool)
        else:
            if str(map_filename).endswith(".npy"):
                self.static_map = np.load(map_filename).astype(bool)
            else:
                self.static_map = load_image_map(map_filename, dims)
        self.origin_coords = np.array([(x_width / 2.) / grid_cell_size, (y_width / 2.) / grid_cell_size])
        self.map = None

    def world_coordinates_to_map_indices(self, pos):
        gx = int(np.floor(self.origin_coords[0] - pos[1] / self.grid_cell_size))
        gy = int(np.floor(self.origin_coords[1] + pos[0] / self.grid_cell_size))
        in_map = 0 <= gx < self.static_map.shape[0] and 0 <= gy < self.static_map.shape[1]
        return np.array([gx, gy]), in_map

    def bits(self):
        """uint32 [rows * ceil(cols/32)] bitmap, bit c&31 of word (r, c>>5) = static_map[r, c]."""
        rows, cols = self.static_map.shape
        wpr = (cols + 31) // 32
        out = np.zeros((rows, wpr), dtype=np.uint32)
        r, c = np.nonzero(self.static_map)
        np.bitwise_or.at(out, (r, c >> 5), (np.uint32(1) << (c & 31).astype(np.uint32)))
        return out.reshape(-1)
