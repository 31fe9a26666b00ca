!! Reference-side plugin: a mergerTreeEvolver that evolves SETS of forests through libglcb200.so.
!!
!! To be added to the Galacticus tree as source/merger_trees/evolver/B200.F90 (next to standard.F90, which it extends).
!! Selected by
!!
!!   <mergerTreeEvolver value="B200"/>
!!
!! It cannot be compiled in the image this repository is developed in (no gfortran/GSL/HDF5/FoX: SURVEY.md 8c).

  !!{RST
  A merger tree evolver which flattens forests, hands them to the batching tree scheduler of ``libglcb200``
  (``glc_forest_evolve``) and rebuilds the surviving nodes from the records that come back.
  !!}

  use, intrinsic :: ISO_C_Binding              , only : c_ptr, c_int32_t, c_int64_t, c_double
  use            :: Node_Evolver_B200_Interface

  !![
  <mergerTreeEvolver name="mergerTreeEvolverB200" docformat="rst">
    <description>
    A merger tree evolver for :galacticus-class:`mergerTreeNodeEvolverB200`. ``mergerTreeEvolverStandard`` walks one tree,
    depth first, and calls the node evolver for one node at a time (``evolver/standard.F90:398-577``, call at ``:452``);
    with a batched node evolver that walk is the serial fraction. This class keeps the rules of the walk---the
    evolvability test (``:723-760``), the limits of ``standardTimeEvolveTo`` (``:762-1035``), ``standardPromote`` and
    ``standardMerge`` (``node_evolver/standard.F90:1241-1356``) and the node operator hooks that go with them---but applies
    them to all nodes of a set of forests at once: ``glc_forest_evolve`` (``galacticus_b200/csrc/host/glc_forest.hpp``)
    gathers, round by round, every node that is allowed to move, evolves them in one batched call, and performs the
    promotions and node mergers on the host.

    Forests never interact during evolution (``tasks/evolve_forests/_class.F90:622,766``), so ``taskEvolveForests`` gives
    each :term:`GPU` (one host thread, one evolver) its own queue of forests.
    </description>
  </mergerTreeEvolver>
  !!]
  type, extends(mergerTreeEvolverStandard) :: mergerTreeEvolverB200
     !!{RST
     Implementation of a merger tree evolver which evolves sets of forests through ``libglcb200``.
     !!}
     private
   contains
     procedure :: evolve => b200ForestEvolve
  end type mergerTreeEvolverB200

contains

  subroutine b200ForestEvolve(self,tree,timeEnd,treeDidEvolve,suspendTree,deadlockReporting,systemClockMaximum,initializationLock,status)
    !!{RST
    Evolves all properties of a merger tree (and of every tree linked to it through ``tree%nextTree``: a forest) to the
    specified time. Same interface as ``standardEvolve`` (``evolver/standard.F90:291``).
    !!}
    use :: Error                       , only : Error_Report            , errorStatusSuccess
    use :: Merger_Tree_Walkers         , only : mergerTreeWalkerAllNodes
    use :: Galacticus_Nodes            , only : nodeComponentBasic      , nodeComponentDarkMatterProfile, nodeComponentSpin
    implicit none
    class           (mergerTreeEvolverB200   ), intent(inout)                   :: self
    type            (mergerTree              ), intent(inout), target           :: tree
    double precision                          , intent(in   )                   :: timeEnd
    logical                                   , intent(  out)                   :: treeDidEvolve       , suspendTree
    logical                                   , intent(in   )                   :: deadlockReporting
    integer         (kind_int8               ), intent(in   ), optional         :: systemClockMaximum
    type            (ompLock                 ), intent(inout), optional         :: initializationLock
    integer                                   , intent(  out), optional         :: status
    integer         (c_int32_t               ), allocatable  , dimension(:  )   :: parent              , flags, state
    real            (c_double                ), allocatable  , dimension(:  )   :: mass                , time , radiusScale, angularMomentum
    real            (c_double                ), allocatable  , dimension(:,:)   :: records
    type            (treeNodeList            ), allocatable  , dimension(:  )   :: nodes
    type            (glcForestCounters       )                                  :: forestCounters
    type            (glcCounters             )                                  :: counters
    type            (mergerTreeWalkerAllNodes)                                  :: walker
    type            (treeNode                ), pointer                         :: node
    class           (nodeComponentBasic      ), pointer                         :: basic
    integer         (c_int64_t               )                                  :: countNodes          , i
    integer                                                                     :: statusCall
    !$GLC attributes unused :: deadlockReporting, systemClockMaximum, initializationLock

    ! Flatten the forest: one entry per node, parents by 0-based index (-1 for the root of a tree). The index of each node in
    ! the flat arrays is kept in nodes(:) so that the records can be scattered back.
    call self%initializeTree(tree,timeEnd)             ! mergerTreeEvolverStandard: nodeTreeInitialize of every operator
    countNodes=b200ForestCount(tree)
    allocate(parent(countNodes),mass(countNodes),time(countNodes),radiusScale(countNodes),angularMomentum(countNodes),nodes(countNodes))
    allocate(records(GLC_NPROP,countNodes),flags(countNodes),state(countNodes))
    walker=mergerTreeWalkerAllNodes(tree,spanForest=.true.)
    i     =0_c_int64_t
    do while (walker%next(node))
       i                  =  i+1_c_int64_t
       nodes          (i)%node => node
       basic              => node %basic()
       mass           (i) =  basic%mass ()
       time           (i) =  basic%time ()
       radiusScale    (i) =  b200RadiusScale    (node)   ! darkMatterProfile%scale()
       angularMomentum(i) =  b200AngularMomentum(node)   ! spin%angularMomentum()
       call node%uniqueIDSet(i)                          ! flat index, looked up for the parent below
    end do
    do i=1_c_int64_t,countNodes
       if (associated(nodes(i)%node%parent)) then
          parent(i)=int(nodes(i)%node%parent%uniqueID()-1_c_int64_t,c_int32_t)
       else
          parent(i)=-1_c_int32_t
       end if
    end do
    ! Evolve. The library returns 0, the warning GLC_WARN_EVOLVE_FAILED (some node evolves failed: the count is in
    ! forestCounters%failed_evolves; the reference's tolerateFailures decides what to do with the tree,
    ! tasks/evolve_forests/_class.F90:887-897), or a negative error.
    statusCall=glc_forest_evolve(b200NodeEvolverHandle(self%mergerTreeNodeEvolver_),countNodes,parent,mass,time,radiusScale,angularMomentum,records,flags,state,forestCounters,counters)
    if (statusCall < 0) call Error_Report('glc_forest_evolve failed'//{introspection:location})
    if (present(status)) then
       status=errorStatusSuccess
       if (statusCall == GLC_WARN_EVOLVE_FAILED) status=errorStatusFail
    end if
    ! Rebuild the tree at the final time: promoted nodes are destroyed (standardPromote moved their components into the
    ! parent: node_evolver/standard.F90:1241-1327), satellites are attached to their hosts (mergerTreeNodeMergerSingleLevelHierarchy),
    ! the records are deserialized into the surviving nodes (mergerTreeNodeEvolverB200: b200NodeScatter).
    call b200ForestRebuild(self,tree,nodes,records,flags,state)
    treeDidEvolve=forestCounters%evolve_calls > 0_c_int64_t
    suspendTree  =.false.
    return
  end subroutine b200ForestEvolve
