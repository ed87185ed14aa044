from .utils import get_point_position_in_the_img, map_points_to_n_imgs, transform_points_to_n_pieces  # noqa: F401
from .camera import Camera, FoVCamera, get_camera_RT  # noqa: F401
