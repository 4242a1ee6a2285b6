from editanything_b200.app import save_input_to_file  # noqa: F401
from editanything_b200.host import HWC3, get_bounding_box, resize_image, resize_points  # noqa: F401
