from .communicator import create_communicator, partition_seeds, owner_of
from .exchange import exchange_extract
