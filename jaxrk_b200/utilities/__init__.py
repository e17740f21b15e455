from .eucldist import eucldist, sqeucldist_simple, sqeucldist_extension
from .distances import dist
