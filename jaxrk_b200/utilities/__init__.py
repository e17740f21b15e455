from .eucldist import eucldist, sqeucldist_simple, sqeucldist_extension
from .distances import dist
from .gram import (rkhs_gram_cdist, rkhs_gram_cdist_ignore_const, rkhs_gram_cdist_unchecked, choose_representer,
                   choose_representer_from_gram, gram_projection)
